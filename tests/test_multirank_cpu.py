"""World-size-2 gloo test of the LazyESN multi-rank host logic (partition of the groups, upload of
the locally needed cells, per-step exchange of the owned cells).  The per-group arithmetic is done
by the CPU oracle here -- the CUDA kernels are covered by the `-m gpu` tests; this test pins the
plumbing that surrounds them when one process drives one GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT, decode_boundary, load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_dir):
    sys.path.insert(0, ROOT)
    import scipy.sparse as sp
    from oracle import esn_oracle as oc
    from xesn_b200 import halo
    from xesn_b200.lazyesn import exchange_owned, partition_groups

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", f"lazy_{name}.npz"))
        field = z["field"]
        nsp = field.ndim - 1
        names = ("x", "y", "z")[:nsp]
        chunks = tuple(int(c) for c in z["chunks"])
        depth = tuple(int(o) for o in z["overlap"])
        b = decode_boundary(z)
        boundary = {names[k]: v for k, v in b.items()} if isinstance(b, dict) else b
        rules = halo.normalise_boundary(boundary, names, depth)
        grid, in_map, out_map, fill = halo.build_maps(field.shape[:-1], chunks, depth, rules)
        G = in_map.shape[0]
        g0, g1 = partition_groups(G, rank, world)
        W = sp.csr_matrix(z["W"])
        Win, bvec, leak = z["Win"], z["b"], float(z["leak"])
        wout = z["wout_groups"].reshape(G, out_map.shape[1], -1)
        n_spinup, n_steps = int(z["n_spinup"]), int(z["n_steps"])
        flat = field.reshape(-1, field.shape[-1])

        def gather(cells, g):
            m = in_map[g]
            u = np.where(m >= 0, cells[np.maximum(m, 0)], fill[-1 - np.minimum(m, -1)])
            return np.nan_to_num(u, nan=0.0)

        # spin-up of my groups on the truth (mask taken from the first sample, lazyesn.py:362-364)
        states = {}
        for g in range(g0, g1):
            r = np.zeros(W.shape[0])
            m = in_map[g]
            first = np.where(m >= 0, flat[np.maximum(m, 0), 0], fill[-1 - np.minimum(m, -1)])
            keep = ~np.isnan(first)
            for t in range(n_spinup):
                u = np.where(m >= 0, flat[np.maximum(m, 0), t], fill[-1 - np.minimum(m, -1)])
                r = oc.leaky_step(r, np.where(keep, u, 0.0), W, Win, bvec, leak)
            states[g] = r

        owned = torch.from_numpy(out_map.reshape(world, -1).astype(np.int64))
        cur = torch.from_numpy(flat[:, n_spinup].copy())
        nxt = cur.clone()
        gathered = torch.empty(tuple(owned.shape), dtype=torch.float64)
        pred = np.zeros((flat.shape[0], n_steps + 1))
        pred[:, 0] = cur.numpy()
        for n in range(1, n_steps + 1):
            nxt.fill_(-12345.0)  # anything I do not own must arrive through the exchange
            cells = cur.numpy()
            for g in range(g0, g1):
                states[g] = oc.leaky_step(states[g], gather(cells, g), W, Win, bvec, leak)
                nxt[torch.from_numpy(out_map[g].astype(np.int64))] = torch.from_numpy(wout[g] @ states[g])
            exchange_owned(nxt, owned, rank, gathered)
            pred[:, n] = nxt.numpy()
            cur, nxt = nxt, cur
        np.save(os.path.join(out_dir, f"pred_{rank}.npy"), pred)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["x_periodic", "xy_periodic", "xy_land"])
def test_two_rank_forecast_matches_reference(tmp_path, name):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    z = load_golden(f"lazy_{name}.npz")
    ref = z["pred"].reshape(-1, z["pred"].shape[-1])
    for rank in range(world):
        got = np.load(tmp_path / f"pred_{rank}.npy")
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)


def test_partition_rules():
    from xesn_b200.lazyesn import partition_groups
    assert partition_groups(16, 0, 1) == (0, 16)
    assert [partition_groups(16, r, 4) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]
    with pytest.raises(ValueError):
        partition_groups(6, 0, 4)


# --------------------------------------------------------------------------- CostFunction: sets dealt to the ranks, costs gathered
def _cost_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from tests.test_host_logic import MACRO_CONFIG
    from xesn_b200 import ESN
    from xesn_b200.cost import CostFunction

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Stub(CostFunction):
            __slots__ = ("seen",)

            def _evaluate_members(self, members):      # the device evaluation of a batch (GPU tests cover it)
                self.seen.extend(m.input_kwargs["factor"] for m in members)
                return np.array([100.0 * m.input_kwargs["factor"] + m.leak_rate for m in members])

        cf = Stub(ESN, None, None, MACRO_CONFIG, max_batch=2)
        cf.seen = []
        sets = np.array([[0.1 * i, 0.5, 0.01 * i, -3.0] for i in range(1, 8)])
        out = cf(sets)
        np.save(os.path.join(out_dir, f"cost_{rank}.npy"), out)
        np.save(os.path.join(out_dir, f"seen_{rank}.npy"), np.array(cf.seen))
    finally:
        dist.destroy_process_group()


def test_two_rank_cost_function(tmp_path):
    """BASELINE configs[4] plumbing: every rank evaluates its share of the parameter sets and every rank ends up with
    all costs, in the order of the sets."""
    world = 2
    mp.spawn(_cost_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = np.array([100.0 * 0.1 * i + 0.01 * i for i in range(1, 8)]).reshape(-1, 1)
    for rank in range(world):
        np.testing.assert_allclose(np.load(tmp_path / f"cost_{rank}.npy"), want, rtol=1e-12)
    np.testing.assert_allclose(np.load(tmp_path / "seen_0.npy"), [0.1, 0.3, 0.5, 0.7], rtol=1e-12)
    np.testing.assert_allclose(np.load(tmp_path / "seen_1.npy"), [0.2, 0.4, 0.6], rtol=1e-12)
