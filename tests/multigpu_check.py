#!/usr/bin/env python
"""Multi-GPU LazyESN parity check (run on a GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Every rank trains/forecasts its slab of groups on its own GPU; the result is compared with the
CPU oracle (test infrastructure) on rank 0."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import esn_oracle as oc  # noqa: E402
from xesn_b200 import LazyESN  # noqa: E402

KW = dict(
    input_kwargs=dict(factor=0.5, distribution="uniform", normalization="svd", random_seed=0),
    adjacency_kwargs=dict(factor=0.9, distribution="uniform", normalization="svd", is_sparse=True,
                          connectedness=5, format="csr", random_seed=1),
    bias_kwargs=dict(factor=0.5, distribution="uniform", random_seed=2),
)


def case(shape, chunks, overlap, boundary, N, rank, engine):
    os.environ["XESN_B200_EXCHANGE"] = engine
    names = ("x", "y")[:len(shape)]
    rs = np.random.RandomState(7)
    field = rs.normal(size=shape + (420,))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        esn = LazyESN(dict(zip(names, chunks)), dict(zip(names, overlap)), boundary, N, 0.6, 1e-4,
                      **{k: dict(v) for k, v in KW.items()})
    esn.build()
    esn.train(field[..., :350], n_spinup=5, batch_size=128)
    wg = esn.Wout_groups                       # collective
    pred = esn.predict(field[..., 350:], n_steps=15, n_spinup=40)
    if rank == 0:
        depth = {a: overlap[a] for a in range(len(shape))}
        depth[len(shape)] = 0
        ob = {names.index(k): v for k, v in boundary.items()} if isinstance(boundary, dict) else boundary  # oracle: by axis
        W = sp.csr_matrix(esn.W)
        ref = oc.lazy_train(field[..., :350], chunks, depth, ob, 5, 128, W, esn.Win, esn.bias_vector, 0.6, 1e-4)
        ref = ref.reshape(wg.shape)
        e1 = oc.max_normalised_err(wg, ref)
        refp = oc.lazy_predict(field[..., 350:], chunks, depth, ob, 15, 40, W, esn.Win, esn.bias_vector, 0.6, wg.reshape(ref.shape).reshape((-1,) + ref.shape[1:]).reshape(tuple(s // c for s, c in zip(shape, chunks)) + ref.shape[1:]))
        # white-noise data makes the closed loop unstable: compare per step, relative to the step's scale
        ax = tuple(range(pred.ndim - 1))
        rel = np.abs(pred - refp).max(axis=ax) / np.maximum(np.abs(refp).max(axis=ax), 1.0)
        e2 = float(rel[:8].max())
        print(f"case {shape} chunks {chunks} engine {engine}: Wout err {e1:.2e}, forecast rel err first 8 steps {e2:.2e}, "
              f"all {float(rel.max()):.2e}", flush=True)
        assert e1 < 1e-6 and e2 < 1e-9, (e1, e2)


def main():
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    # engines: dataflow kernel with neighbour-only P2P stores (default), persistent kernel with every cell to every
    # rank, per-step launches + NCCL
    for engine in ("flow", "p2p", "nccl"):
        case((64,), (4,), (1,), "periodic", 300, rank, engine)
        case((8 * max(world // 2, 1), 12), (2, 3), (1, 1), "periodic", 200, rank, engine)
        case((16,), (2,), (2,), 0.0, 150, rank, engine)
        case((32, 6), (4, 3), (2, 1), {"x": "reflect", "y": "periodic"}, 150, rank, engine)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(f"MULTIGPU OK (world {world})")


if __name__ == "__main__":
    main()
