"""CPU tests of the host side: option handling mirrors the reference's API tests
(xesn/test/esn.py, xesn/test/lazy.py), halo index maps against the oracle's np.pad halo,
the C-ABI library loads and exports every declared symbol.  No GPU compute here."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest

from oracle import esn_oracle as oc
from tests.conftest import ESN_KW, ROOT
from xesn_b200 import ESN, LazyESN, XesnError, _lib
from xesn_b200 import halo
from xesn_b200.esn import auto_chunk


def kw():
    return {k: dict(v) for k, v in ESN_KW.items()}


# --------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python __graft_entry__.py` first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "xesn_b200.h")).read()
    declared = set(re.findall(r"\b(xesn_[a-z0-9_]+)\s*\(", header))
    declared -= {"xesn_series"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/xesn_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib.xesn_abi_version.restype = ctypes.c_int
    assert lib.xesn_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    esn = ESN(3, 3, 20, 0.5, 1e-6, **kw())
    esn.build()
    with pytest.raises(XesnError):
        esn.train(np.zeros((3, 10)))
    with pytest.raises(XesnError):
        ESN(3, 3, 20, 0.5, 1e-6, **kw()).train(np.zeros((3, 10)))  # not built


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "xesn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", ""), f"{f} mentions the oracle package"


# --------------------------------------------------------------------------- ESN options (xesn/test/esn.py)
def test_defaults_warn_and_filtering():
    with pytest.warns(UserWarning, match="Did not find 'input_kwargs'"):
        esn = ESN(3, 3, 20, 0.5, 1e-6, adjacency_kwargs=kw()["adjacency_kwargs"], bias_kwargs=kw()["bias_kwargs"])
    assert esn.input_kwargs == {"factor": 1.0, "distribution": "uniform", "normalization": "multiply",
                                "is_sparse": False, "random_seed": None}
    with pytest.warns(UserWarning, match="'blah' not allowed"):
        k = kw()
        k["bias_kwargs"]["blah"] = 1
        esn = ESN(3, 3, 20, 0.5, 1e-6, **k)
    assert "blah" not in esn.bias_kwargs
    # user dicts are filtered, NOT merged with the defaults (esn.py:375-387)
    k = kw()
    k["adjacency_kwargs"] = {"factor": 0.9, "distribution": "uniform", "is_sparse": True, "connectedness": 3}
    esn = ESN(3, 3, 20, 0.5, 1e-6, **k)
    assert "format" not in esn.adjacency_kwargs
    esn.build()
    assert esn.W.format == "coo"  # SparseRandomMatrix default (matrix.py:240)
    assert esn.input_factor == 0.5 and esn.adjacency_factor == 0.9 and esn.bias_factor == 0.5


def test_negative_bias_and_density_errors():
    k = kw()
    k["bias_kwargs"]["factor"] = -1.0
    with pytest.raises(ValueError):
        ESN(3, 3, 20, 0.5, 1e-6, **k)
    k = kw()
    k["adjacency_kwargs"].pop("connectedness")
    k["adjacency_kwargs"]["density"] = 1.5
    with pytest.raises(ValueError):
        ESN(3, 3, 20, 0.5, 1e-6, **k).build()
    k["adjacency_kwargs"]["density"] = 0.5
    with pytest.warns(RuntimeWarning, match="density is >20%"):
        ESN(3, 3, 20, 0.5, 1e-6, **k).build()


def test_build_shapes_and_str():
    esn = ESN(4, 2, 30, 0.5, 1e-6, **kw())
    assert str(esn) == repr(esn) and str(esn).startswith("ESN\n")
    esn.build()
    assert esn.W.shape == (30, 30) and esn.W.nnz == 150
    assert esn.Win.shape == (30, 4) and esn.bias_vector.shape == (30,)
    with pytest.raises(Exception, match="Wout has not been computed"):
        esn.to_xds()
    attrs = esn._get_attrs()
    assert list(attrs) == ["n_input", "n_output", "n_reservoir", "leak_rate", "tikhonov_parameter",
                           "input_kwargs", "adjacency_kwargs", "bias_kwargs"]


def test_auto_chunk_bounds():
    assert auto_chunk(16000, 1, 20000, None) >= 512
    assert auto_chunk(16000, 1, 20000, 100) == 100
    assert auto_chunk(100, 1, 10, None) == 10
    assert auto_chunk(6000, 256, 4096, None) >= 64
    assert 48 * 6000 * 256 * auto_chunk(6000, 256, 4096, None) <= (12 << 30)      # the scratch budget (12 of the 180 GB)


# --------------------------------------------------------------------------- LazyESN shell (xesn/test/lazy.py)
def lazy_kw(**over):
    d = dict(esn_chunks={"x": 3, "time": 1000}, overlap={"x": 1, "time": 0}, boundary=0, n_reservoir=20,
             leak_rate=0.5, tikhonov_parameter=1e-6, persist=True, **kw())
    d.update(over)
    return d


def test_lazy_basic_properties():
    esn = LazyESN(**lazy_kw())
    assert str(esn) == repr(esn) and str(esn).startswith("LazyESN\n")
    assert esn.esn_chunks == esn.output_chunks
    assert esn.input_chunks == {"x": 5, "time": 1000}
    assert esn.n_input == 5 and esn.n_output == 3
    assert esn.ndim_state == 1
    assert esn._r_chunks == (20,)
    assert esn._Wout_chunks == (3, 20)
    e3 = LazyESN(**lazy_kw(esn_chunks={"x": 2, "y": 3, "z": 5}, overlap={"x": 1, "y": 1, "z": 0}))
    assert e3.ndim_state == 3 and e3._r_chunks == (1, 1, 20) and e3._Wout_chunks == (30, 20, 1)
    assert e3.n_input == 4 * 5 * 5 and e3.n_output == 30


@pytest.mark.parametrize("attr,expected", [("overlap", 0), ("esn_chunks", -1)])
def test_lazy_default_time(attr, expected):
    k = lazy_kw()
    k[attr].pop("time")
    esn = LazyESN(**k)
    assert getattr(esn, attr)["time"] == expected


def test_lazy_errors():
    with pytest.raises(NotImplementedError):
        LazyESN(**lazy_kw(overlap={"x": 1, "y": 1, "z": 1, "time": 0},
                          esn_chunks={"x": 2, "y": 2, "z": 2, "time": 1000}))
    with pytest.raises(ValueError):
        LazyESN(**lazy_kw(esn_chunks={"x": 3, "y": -1, "time": 1000}, overlap={"x": 3, "y": 0, "time": 1000}))
    attrs = LazyESN(**lazy_kw())._get_attrs()
    assert list(attrs)[:7] == ["esn_chunks", "overlap", "boundary", "n_reservoir", "leak_rate",
                               "tikhonov_parameter", "persist"]


# --------------------------------------------------------------------------- halo maps == dask overlap restated
CASES = [
    ((12,), (3,), (1,), "periodic"),
    ((4,), (2,), (1,), 0.0),
    ((9,), (3,), (2,), float("nan")),
    ((6, 8), (3, 4), (1, 1), "periodic"),
    ((6, 8), (3, 2), (2, 1), {0: "reflect", 1: 0.5}),
    ((6, 4), (2, 4), (1, 0), "periodic"),
    ((4, 6, 2), (2, 3, 2), (1, 1, 0), {0: "periodic", 1: 0.0, 2: "reflect"}),
    ((4, 6, 4), (2, 3, 2), (1, 1, 0), {0: 7.0, 1: 3.0}),
]


@pytest.mark.parametrize("shape,chunks,ovl,boundary", CASES)
def test_halo_maps_match_oracle_blocks(shape, chunks, ovl, boundary):
    nsp = len(shape)
    names = ("x", "y", "z")[:nsp]
    b = {names[k]: v for k, v in boundary.items()} if isinstance(boundary, dict) else boundary
    rules = halo.normalise_boundary(b, names, ovl)
    grid, in_map, out_map, fill = halo.build_maps(shape, chunks, ovl, rules)
    field = np.random.RandomState(1).normal(size=shape + (3,))
    depth = {a: ovl[a] for a in range(nsp)}
    depth[nsp] = 0
    flat = field.reshape(-1, 3)
    blocks = list(oc.halo_blocks(field, chunks, depth, boundary))
    assert len(blocks) == in_map.shape[0] == int(np.prod(grid))
    for k, (g, blk) in enumerate(blocks):
        want = blk.reshape(-1, 3)
        got = np.where(in_map[k][:, None] >= 0, flat[np.maximum(in_map[k], 0)], fill[-1 - np.minimum(in_map[k], -1)][:, None])
        np.testing.assert_array_equal(got, np.nan_to_num(want, nan=0.0))
        inner = oc.inner_mask(blk.shape, depth).reshape(-1)
        np.testing.assert_array_equal(flat[out_map[k]], want[inner])
    # every cell is predicted by exactly one group
    assert sorted(out_map.reshape(-1).tolist()) == list(range(int(np.prod(shape))))


def test_halo_errors():
    with pytest.raises(ValueError):
        halo.build_maps((10,), (3,), (1,), ["periodic"])
    with pytest.raises(ValueError):
        halo.normalise_boundary({"y": 0.0}, ("x",), (1,))
    with pytest.raises(NotImplementedError):
        halo.normalise_boundary("nearest", ("x",), (1,))


@pytest.mark.parametrize("grid", [(4,), (2, 3), (2, 3, 2)])
def test_wout_block_layout_roundtrip(grid):
    G, O, N = int(np.prod(grid)), 5, 7
    w = np.random.RandomState(0).normal(size=(G, O, N))
    blocks = halo.wout_to_blocks(w, grid)
    np.testing.assert_array_equal(blocks, oc.wout_block_layout(w.reshape(grid + (O, N))))
    np.testing.assert_array_equal(halo.blocks_to_wout(blocks, grid, O, N), w)


# --------------------------------------------------------------------------- ESNEnsemble host logic (no GPU needed)
def _member(seed_adj=1, n_res=80, n_in=3):
    from xesn_b200 import ESN
    kw = {k: dict(v) for k, v in ESN_KW.items()}
    kw["adjacency_kwargs"]["random_seed"] = seed_adj
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ESN(n_in, n_in, n_res, 0.5, 1e-4, **kw)


def test_ensemble_rejects_mismatched_members():
    """Members must have the same sizes (constructor) and the same sparsity pattern of W (build, checked on the
    host before any device call): the batch shares one CSR structure, xesn_reservoir_set_group_params."""
    from xesn_b200 import ESNEnsemble, XesnError
    with pytest.raises(AssertionError):
        ESNEnsemble([_member(n_res=80), _member(n_res=90)])
    with pytest.raises(AssertionError):
        ESNEnsemble([])
    ens = ESNEnsemble([_member(seed_adj=1), _member(seed_adj=7)])
    with pytest.raises(XesnError, match="sparsity pattern"):
        ens.build()
    with pytest.raises(XesnError, match="build"):
        ESNEnsemble([_member()]).train(np.zeros((3, 10)))


def test_bench_workloads_are_the_baseline_configs():
    """bench.py names one workload per BASELINE.json configuration; unit-steps follow SURVEY.md section 8(d)."""
    import bench
    w = bench.WORKLOADS
    assert w["eager16k"]["n_reservoir"] == 16000 and w["eager1k"]["n_reservoir"] == 1000
    assert (w["lazy"]["chunk"], w["lazy"]["overlap"], w["lazy"]["n_reservoir"]) == (16, 1, 2000)
    assert (w["lazy2d"]["chunk"], w["lazy2d"]["overlap"], w["lazy2d"]["n_reservoir"]) == (32, 2, 6000)
    assert w["lazy2d"]["groups_per_gpu"] == w["lazy2d"]["group_rows_per_gpu"] * w["lazy2d"]["group_cols"]
    assert w["macro"]["groups_per_gpu"] * 8 == 64 and w["macro"]["n_reservoir"] == 4000
    assert bench.unit_steps(w["eager16k"], 1) == 16000 * (20000 + 1500)
    assert bench.unit_steps(w["macro"], 8) == 8 * 4000 * (20000 + 5 * 600)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port on the host cores) prints ONE JSON line with the keys the driver
    reads; run here on the reference's own CPU-sized configuration (BASELINE configs[0]) and on a multi-reservoir one."""
    import json
    import subprocess
    import sys
    for workload in ("eager1k", "lazy"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                              "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == "unit-steps/s" and d["higher_is_better"] is True
        assert d["metric"].startswith("reservoir unit-steps/sec") and d["value"] > 0 and d["vs_baseline"] is None
        assert d["config"]["workload"] == workload and d["dtype"] == "f64" and d["data"] == "synthetic"
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle port" in cb["sample"]
        assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_ensemble_host_build_reuses_unscaled_matrices():
    """ESNEnsemble.build_host draws every random matrix once per seed and applies each member's factor with the
    reference's expression (matrix.py:183-193, :296-316): same W, Win, bias as building the members one by one."""
    from xesn_b200 import ESNEnsemble
    rs = np.random.RandomState(5)

    def members():
        out = []
        for f_in, f_adj, f_b in rs.uniform(0.01, 2.0, size=(4, 3)):
            kw = {k: dict(v) for k, v in ESN_KW.items()}
            kw["input_kwargs"]["factor"], kw["adjacency_kwargs"]["factor"], kw["bias_kwargs"]["factor"] = f_in, f_adj, f_b
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out.append(ESN(7, 7, 300, 0.3, 1e-5, **kw))
        return out

    state = rs.get_state()
    fast = members()
    rs.set_state(state)
    slow = members()
    ESNEnsemble(fast).build_host()
    for a, b in zip(fast, slow):
        b.build()
        assert (a.W != b.W).nnz == 0 or np.allclose(a.W.toarray(), b.W.toarray(), rtol=1e-13, atol=0)
        assert np.array_equal(a.W.tocsr().indices, b.W.tocsr().indices)
        np.testing.assert_allclose(a.W.toarray(), b.W.toarray(), rtol=1e-13, atol=0)   # ARPACK jitter in sigma_max only
        np.testing.assert_allclose(a.Win, b.Win, rtol=1e-15, atol=0)
        np.testing.assert_array_equal(a.bias_vector, b.bias_vector)


# --------------------------------------------------------------------------- CostFunction host logic
MACRO_CONFIG = {
    "esn": dict(n_input=3, n_output=3, n_reservoir=60, leak_rate=0.5, tikhonov_parameter=1e-6,
                **{k: dict(v) for k, v in ESN_KW.items()}),
    "training": {"n_spinup": 5, "batch_size": 50},
    "macro_training": {
        "parameters": {"input_factor": [0.0, 2.0], "adjacency_factor": [0.0, 2.0], "leak_rate": [0.0, 1.0],
                       "tikhonov_parameter": [1e-12, 1.0]},
        "transformations": {"tikhonov_parameter": "log10"},
        "forecast": {"n_spinup": 10, "n_samples": 2, "n_steps": 20},
        "cost_upper_bound": 1e9,
        "cost_terms": {"nrmse": 1.0, "psd_nrmse": 1.0},
    },
}


def test_transform_known_answers():
    """The docstring examples of the reference (xesn/optim.py:64-70 and :116-123)."""
    from xesn_b200.cost import inverse_transform, transform
    t = {"input_factor": "log", "adjacency_factor": "log10"}
    assert transform({"input_factor": 0.5, "adjacency_factor": 0.5, "bias_factor": 0.5}, t) == {
        "input_factor": -0.6931471805599453, "adjacency_factor": -0.3010299956639812, "bias_factor": 0.5}
    got = inverse_transform({"input_factor": -0.69, "adjacency_factor": -0.3, "bias_factor": 0.5}, t)
    assert got == {"input_factor": 0.5015760690660556, "adjacency_factor": 0.5011872336272722, "bias_factor": 0.5}
    assert inverse_transform({"a": [0.0, 1.0]}, {"a": "log10"}) == {"a": [1.0, 10.0]}
    with pytest.raises(KeyError):
        inverse_transform({"a": 1.0}, {"b": "log"})
    with pytest.raises(NotImplementedError):
        transform({"a": 1.0}, {"a": "sqrt"})


def test_update_esn_kwargs_and_candidates():
    """Reference xesn/utils.py:65-100 and cost.py:177-190: factors go into the *_kwargs dicts, the rest to the top level."""
    from xesn_b200.cost import CostFunction, update_esn_kwargs
    orig = {"leak_rate": 0.5, "input_kwargs": {"factor": 1.0, "random_seed": 0}}
    new = update_esn_kwargs({"input_factor": 0.25, "bias_factor": 0.1, "leak_rate": 0.9, "n_reservoir": 10}, orig)
    assert new == {"leak_rate": 0.9, "input_kwargs": {"factor": 0.25, "random_seed": 0}, "bias_kwargs": {"factor": 0.1},
                   "n_reservoir": 10}
    assert orig["input_kwargs"]["factor"] == 1.0                      # the original is not modified
    cf = CostFunction(ESN, None, None, MACRO_CONFIG)
    kws = cf.candidate_kwargs([[0.3, 0.8, 0.4, -3.0], [1.5, 0.2, 0.9, -8.0]])
    assert [k["input_kwargs"]["factor"] for k in kws] == [0.3, 1.5]
    assert [k["adjacency_kwargs"]["factor"] for k in kws] == [0.8, 0.2]
    assert [k["leak_rate"] for k in kws] == [0.4, 0.9]
    np.testing.assert_allclose([k["tikhonov_parameter"] for k in kws], [1e-3, 1e-8], rtol=1e-14)
    assert cf.candidate_kwargs([0.3, 0.8, 0.4, 1e-3], is_transformed=False)[0]["tikhonov_parameter"] == 1e-3
    assert kws[0]["bias_kwargs"] == MACRO_CONFIG["esn"]["bias_kwargs"]
    assert cf.batch_size(4000, 64) >= 8 and cf.batch_size(16000, 64) >= 1
    with pytest.raises(NotImplementedError):
        CostFunction(ESN, None, None, {"macro_training": {"cost_terms": {"mae": 1.0}}})
    assert CostFunction(LazyESN, None, None, MACRO_CONFIG).ESN is LazyESN   # LazyESN candidates are accepted since round 2


def test_cost_function_batches_and_orders_the_sets():
    """__call__ builds one member per set, evaluates them in batches of `max_batch` and returns (n_sets, 1) in the
    order of the sets (cost.py:94-122); the per-batch device evaluation is stubbed here (it is covered by the GPU
    test of ESNEnsemble.cost)."""
    from xesn_b200.cost import CostFunction, shard

    class Stub(CostFunction):
        __slots__ = ("batches",)

        def _evaluate_members(self, members):
            self.batches.append(len(members))
            return np.array([m.leak_rate + 10.0 * m.input_kwargs["factor"] for m in members])

    cf = Stub(ESN, None, None, MACRO_CONFIG, max_batch=3)
    cf.batches = []
    sets = np.array([[0.1 * i, 0.5, 0.01 * i, -3.0] for i in range(1, 8)])
    out = cf(sets)
    assert out.shape == (7, 1) and cf.batches == [3, 3, 1]
    np.testing.assert_allclose(out[:, 0], [0.01 * i + 1.0 * i for i in range(1, 8)], rtol=1e-12)
    assert shard(7, 0, 2) == [0, 2, 4, 6] and shard(7, 1, 2) == [1, 3, 5] and shard(2, 3, 4) == []


def test_cost_function_device_glue_calls_match_the_ensemble_api(monkeypatch):
    """`CostFunction._evaluate_members` is a thin sequence of ESNEnsemble calls (each covered by the GPU tests); here
    the calls are bound against the real signatures with a recording stand-in, so that a keyword mismatch cannot hide
    behind the missing GPU."""
    import inspect
    from xesn_b200 import cost as cost_mod
    from xesn_b200.ensemble import ESNEnsemble
    calls = []

    class Fake:
        def __init__(self, members):
            inspect.signature(ESNEnsemble.__init__).bind(self, members)
            self.members = list(members)

        def build(self):
            calls.append("build")

        def train(self, *a, **k):
            inspect.signature(ESNEnsemble.train).bind(self, *a, **k)
            calls.append(("train", k))
            return self

        def cost(self, *a, **k):
            bound = inspect.signature(ESNEnsemble.cost).bind(self, *a, **k)
            calls.append(("cost", dict(bound.arguments)))
            return np.arange(len(self.members), dtype=np.float64), None

    monkeypatch.setattr(cost_mod, "ESNEnsemble", Fake)
    train, windows = np.zeros((3, 100)), [np.zeros((3, 31)), np.zeros((3, 31))]
    cf = cost_mod.CostFunction(ESN, train, windows, MACRO_CONFIG)
    out = cf([[0.3, 0.8, 0.4, -3.0], [1.5, 0.2, 0.9, -8.0], [1.0, 1.0, 0.5, -6.0]])
    np.testing.assert_array_equal(out[:, 0], [0.0, 1.0, 2.0])
    assert calls[0] == "build" and calls[1] == ("train", {"n_spinup": 5, "batch_size": 50})
    args = calls[2][1]
    assert args["n_steps"] == 20 and args["n_spinup"] == 10 and args["cost_upper_bound"] == 1e9
    assert args["cost_terms"] == {"nrmse": 1.0, "psd_nrmse": 1.0} and args["windows"] is windows


# --------------------------------------------------------------------------- dataflow forecast tables (neighbour-only exchange)
@pytest.mark.parametrize("shape,chunks,depth,boundary,world", [
    ((64,), (4,), (1,), "periodic", 4),
    ((64,), (4,), (2,), 0.0, 2),
    ((16, 12), (2, 3), (1, 1), "periodic", 4),
    ((16, 12), (4, 3), (2, 1), {"x": "reflect", "y": "periodic"}, 2),
    ((32,), (4,), (1,), "periodic", 1),
])
def test_flow_tables_route_every_input(shape, chunks, depth, boundary, world):
    """The tables of the dataflow forecast kernel: emulate one exchange step on the host -- every rank fills its local
    ring slots, pushes its exports, and then must find, through `src`, exactly the values a whole-field gather
    through in_map gives (what dask's overlap does in xesn/lazyesn.py:235); imports are neighbour-only."""
    from xesn_b200 import halo
    names = ("x", "y")[:len(shape)]
    rules = halo.normalise_boundary(boundary, names, depth)
    grid, in_map, out_map, fill = halo.build_maps(shape, chunks, depth, rules)
    n_cells = int(np.prod(shape))
    G_all, O = out_map.shape
    per = G_all // world
    value = np.random.RandomState(1).normal(size=n_cells)          # the freshly predicted field
    tabs = [halo.flow_tables(in_map, out_map, n_cells, world, r) for r in range(world)]
    rings = []
    for r in range(world):
        ring = np.full(per * O + int(tabs[r]["n_import_all"][r]), np.nan)
        ring[:per * O] = value[out_map[r * per:(r + 1) * per].reshape(-1)]
        rings.append(ring)
    for r in range(world):
        t = tabs[r]
        assert len(t["import_cell"]) == t["n_import_all"][r]
        for s in range(per * O):
            for e in range(t["exp_ptr"][s], t["exp_ptr"][s + 1]):
                peer, slot = int(t["exp_peer"][e]), int(t["exp_slot"][e])
                assert peer != r and np.isnan(rings[peer][slot])       # exactly one producer per import slot
                rings[peer][slot] = rings[r][s]
    total_imports = 0
    for r in range(world):
        t = tabs[r]
        assert not np.isnan(rings[r]).any()                            # every import slot was served
        np.testing.assert_array_equal(rings[r][per * O:], value[t["import_cell"]])
        src = t["src"].astype(np.int64)
        m = in_map[r * per:(r + 1) * per].astype(np.int64)
        direct = np.where(m >= 0, value[np.maximum(m, 0)], fill[np.maximum(-1 - m, 0)])
        via = np.where(src >= 0, rings[r][np.maximum(src, 0)], fill[np.maximum(-1 - np.maximum(src, -len(fill)), 0)])
        assert (src != halo.SRC_STATIC).all()                          # chunks tile the domain: nothing is static
        np.testing.assert_array_equal(via, direct)
        total_imports += len(t["import_cell"])
        # neighbour-only: a slab imports from at most two other ranks along the split axis
        assert len(set(t["import_owner"].tolist())) <= min(2, world - 1)
    if world == 1:
        assert total_imports == 0
    else:
        assert 0 < total_imports < n_cells                             # halo strips, not the whole field


# --------------------------------------------------------------------------- checkpoint round trip (io.py)
def test_checkpoint_roundtrip_npz_and_zarr_gate(tmp_path):
    """`to_xds` content (readout + constructor kwargs incl. seeds) through the dependency-free .npz twin: the rebuilt
    model has the same matrices (the seeds are the checkpoint, xesn/esn.py:309-338, io.py:13-53) and readout; LazyESN
    readouts travel in the reference's block layout with the singleton axis re-appended (io.py:46-49)."""
    import warnings
    import xesn_b200 as xb
    from xesn_b200 import halo
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        esn = xb.ESN(3, 3, 60, 0.4, 1e-5, **{k: dict(v) for k, v in ESN_KW.items()})
        esn.build()
        esn.Wout = np.random.RandomState(0).normal(size=(3, 60))
        path = str(tmp_path / "eager.npz")
        xb.to_npz(esn, path)
        back = xb.from_npz(path)
        assert type(back) is xb.ESN and back.leak_rate == 0.4 and back.adjacency_kwargs == esn.adjacency_kwargs
        np.testing.assert_allclose(back.Win, esn.Win, rtol=1e-12)
        np.testing.assert_allclose(back.W.toarray(), esn.W.toarray(), rtol=1e-12)
        np.testing.assert_array_equal(back.Wout, esn.Wout)

        lazy = xb.LazyESN({"x": 2, "y": 3, "z": 2}, {"x": 1, "y": 1, "z": 0}, {"x": "periodic", "y": 0.0}, 30, 0.5, 1e-6,
                          **{k: dict(v) for k, v in ESN_KW.items()})
        lazy.build()
        grid = (4, 2, 1)
        wg = np.random.RandomState(1).normal(size=(8, lazy.n_output, 30))
        lazy._grid = grid
        lazy.Wout = halo.wout_to_blocks(wg, grid)
        assert lazy.Wout.shape == (4 * 12, 2 * 30, 1)
        path = str(tmp_path / "lazy.npz")
        xb.to_npz(lazy, path)
        assert np.load(path)["Wout"].shape == (48, 60)          # stored squeezed, like to_xds
        back = xb.from_npz(path)
        assert type(back) is xb.LazyESN and back.boundary == {"x": "periodic", "y": 0.0} and back.overlap["z"] == 0
        np.testing.assert_array_equal(back.Wout, lazy.Wout)
        np.testing.assert_array_equal(halo.blocks_to_wout(back.Wout, back._grid, 12, 30), wg)
        untrained = xb.ESN(3, 3, 60, 0.4, 1e-5, **{k: dict(v) for k, v in ESN_KW.items()})
        with pytest.raises(Exception):
            xb.to_npz(untrained, str(tmp_path / "no.npz"))
    try:
        import xarray  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            xb.from_zarr(str(tmp_path / "store.zarr"))


def test_psd_bin_tables_match_binned_statistic():
    """The annulus cell lists the device kernel uses against scipy.stats.binned_statistic on the same wavenumbers
    (xesn/psd.py:42-65)."""
    from scipy.fft import fftfreq
    from scipy.stats import binned_statistic
    from xesn_b200.psd import bin_tables
    for n_x, nd in ((12, 1), (13, 1), (8, 2), (9, 2), (32, 2)):
        bin_ptr, pix_idx, factor, k_vals = bin_tables(n_x, nd)
        k = n_x * fftfreq(n_x)
        k_tot = k if nd == 1 else np.sqrt(sum(a ** 2 for a in np.meshgrid(k, k))).reshape(-1)
        k_bins = np.arange(.5, n_x // 2 + 1, 1.)
        values = np.random.RandomState(n_x).rand(k_tot.size)
        with np.errstate(all="ignore"):
            ref, _, _ = binned_statistic(k_tot, values, statistic="mean", bins=k_bins)
        got = np.array([values[pix_idx[bin_ptr[b]:bin_ptr[b + 1]]].mean() if bin_ptr[b + 1] > bin_ptr[b] else np.nan
                        for b in range(len(factor))])
        np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(ref), rtol=1e-13)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        np.testing.assert_allclose(factor, np.pi * (k_bins[1:] ** 2 - k_bins[:-1] ** 2))


def _gather_wavefronts(indptr, indices, slot, upc, zero_slot):
    """Independent count of the shared-memory wavefronts of the row gathers (8-byte loads: a half-warp at a time, one
    wavefront per distinct address that shares a 16-wide bank pair with another): 8 rounds per warp plus rounds of 4
    while the warp's longest row lasts -- the access pattern of recur_cluster.cuh."""
    n = len(indptr) - 1
    total = n_sets = 0
    for u0 in range(0, n, upc):
        u1 = min(n, u0 + upc)
        for w0 in range(u0, u1, 32):
            rows = range(w0, min(w0 + 32, u1))
            longest = max(indptr[r + 1] - indptr[r] for r in rows)
            rounds = 8 + (0 if longest <= 8 else -(-(longest - 8) // 4) * 4)
            for h0 in range(w0, min(w0 + 32, u1), 16):
                half = range(h0, min(h0 + 16, u1))
                for k in range(rounds):
                    cols = {indices[indptr[r] + k] for r in half if k < indptr[r + 1] - indptr[r]}
                    if not cols:
                        continue
                    addr = {slot[c] for c in cols}
                    if any(k >= indptr[r + 1] - indptr[r] for r in half):
                        addr.add(zero_slot)
                    total += np.bincount([a % 16 for a in addr], minlength=16).max()
                    n_sets += 1
    return total, n_sets


@pytest.mark.parametrize("n,team", [(1000, 8), (2000, 4), (333, 2)])
def test_bank_layout_is_a_permutation_and_removes_conflicts(n, team):
    """xesn_bank_layout (host code of the library): the placement is a permutation into the slots, the reported
    wavefront counts agree with an independent count, and the search removes most of the bank conflicts of the natural
    placement (conflict-free = one wavefront per non-empty half-warp gather)."""
    import scipy.sparse as sp
    from xesn_b200 import SparseRandomMatrix
    W = sp.csr_matrix(SparseRandomMatrix(n_rows=n, n_cols=n, factor=0.9, distribution="uniform", normalization="multiply",
                                         connectedness=5, format="csr", random_seed=1)())
    indptr, indices = W.indptr.astype(np.int32), W.indices.astype(np.int32)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.xesn_bank_layout.restype = ctypes.c_int64
    lib.xesn_bank_layout.argtypes = _lib.SIGNATURES["xesn_bank_layout"][1]
    upc = ((n + team - 1) // team + 7) & ~7
    n_slots = (n + 15) // 16 * 16 + 32
    slot = np.zeros(n, np.int32)
    natural = ctypes.c_int64()
    after = lib.xesn_bank_layout(n, indptr.ctypes.data, indices.ctypes.data, upc, n_slots, slot.ctypes.data,
                                 ctypes.byref(natural))
    assert after > 0
    assert len(np.unique(slot)) == n and slot.min() >= 0 and slot.max() < n_slots
    # a warp's states stay inside the warp's own run of 32 slots (remote stores of a warp: one 256-byte segment)
    for u0 in range(0, n, upc):
        for w0 in range(u0, min(n, u0 + upc), 32):
            w1 = min(n, u0 + upc, w0 + 32)
            assert sorted(slot[w0:w1]) == list(range(w0, w1))
    assert _gather_wavefronts(indptr, indices, np.arange(n), upc, n_slots + 1)[0] == natural.value
    count, floor = _gather_wavefronts(indptr, indices, slot, upc, n_slots + 1)   # floor: one wavefront per gather
    assert count == after
    assert after <= natural.value
    assert after - floor <= 0.6 * (natural.value - floor), (floor, after, natural.value)
    # bad arguments are refused
    assert lib.xesn_bank_layout(n, indptr.ctypes.data, indices.ctypes.data, upc, n_slots + 1, slot.ctypes.data, None) < 0


def test_gram_and_readout_interleaved_layout():
    """The Gram matrices and readouts of a training call are views of ONE array (G, N + O, N) when O <= 256 -- the layout the
    library recognises by `wout == gram + N*N` (include/xesn_b200.h) and factors with the right-hand sides as extra rows --
    and two separate arrays otherwise."""
    import torch
    from xesn_b200.esn import gram_and_readout
    G, N, O = 3, 20, 5
    gram, wout, both = gram_and_readout(G, N, O, "cpu")
    assert both is not None and both.shape == (G, N + O, N)
    assert gram.shape == (G, N, N) and wout.shape == (G, O, N)
    assert wout.data_ptr() == gram.data_ptr() + 8 * N * N
    assert gram.stride() == ((N + O) * N, N, 1) and wout.stride() == ((N + O) * N, N, 1)
    assert gram[1].data_ptr() == both[1].data_ptr() and wout[2].data_ptr() == both[2, N:].data_ptr()
    both.zero_()
    wout[1, 2, 3] = 7.0
    assert both[1, N + 2, 3] == 7.0 and float(gram.abs().sum()) == 0.0
    gram2, wout2, both2 = gram_and_readout(2, 10, 300, "cpu")
    assert both2 is None and gram2.is_contiguous() and wout2.is_contiguous()
    assert wout2.data_ptr() != gram2.data_ptr() + 8 * 10 * 10
