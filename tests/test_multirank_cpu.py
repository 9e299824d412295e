"""N>1 host logic on CPU (gloo, world_size 2): cell-range partition, dof ownership, local renumbering
and the column-block merge reproduce the single-rank matrix.  The per-rank assembly is done by the
oracle here (no GPU in this container); on the GPU box the same partition feeds libgrmp_cuda
(bench.py --gpus N, tests/test_gpu_parity.py::test_partitioned_assembly_matches_global)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import grmp_b200 as G
import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _assemble_oracle(grid, space):
    A = O.OracleMatrix(space.ndofs, space.ndofs)
    O.blf_assemble(A, grid, space, space, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC)
    return A.csc()


def _worker(rank, world, port, level, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = G.perturb_interior_nodes(G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), level))
    s = G.FESpace(G.H1P2(1, 3), g)
    lp = G.partition.partition(s, rank, world)
    cp, rv, nz = _assemble_oracle(lp.grid, lp.space)
    gcols, bcp, brv, bnz = G.partition.owned_block_to_global(lp, cp, rv, nz)
    # exchange: every rank learns the per-rank owned nnz (the tiny all-gather of the symbolic pass)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([brv.size], dtype=torch.int64))
    blocks = [None] * world
    dist.all_gather_object(blocks, (gcols, bcp, brv, bnz))
    if rank == 0:
        assert [int(c) for c in counts] == [b[2].size for b in blocks]
        colptr, rowval, nzval = G.partition.merge_owned_columns(s.ndofs, blocks)
        np.savez(out, colptr=colptr, rowval=rowval, nzval=nzval, halo=[lp.grid.ncells, g.ncells])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_reproduces_global_matrix(tmp_path):
    level, world = 1, 2
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), level, out), nprocs=world, join=True)
    d = np.load(out)
    g = G.perturb_interior_nodes(G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), level))
    s = G.FESpace(G.H1P2(1, 3), g)
    cp, rv, nz = _assemble_oracle(g, s)
    assert np.array_equal(d["colptr"], cp) and np.array_equal(d["rowval"], rv)
    # owner-computes: every column is complete on its owner, cells are visited in the same relative order
    assert np.array_equal(d["nzval"], nz)


def test_every_dof_has_exactly_one_owner_and_owned_columns_are_complete():
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), 2)
    s = G.FESpace(G.H1P2(1, 3), g)
    world = 3
    owner = G.partition.dof_owner(s, world)
    assert owner.min() == 0 and owner.max() == world - 1
    total = 0
    for r in range(world):
        lp = G.partition.partition(s, r, world)
        total += lp.n_owned
        owned_global = lp.local2global[: lp.n_owned]
        assert np.all(owner[owned_global] == r)
        # completeness: all cells of the global grid touching an owned dof are present locally
        touching = np.nonzero(np.isin(s.celldofs.astype(np.int64) - 1, owned_global).any(axis=1))[0]
        assert np.array_equal(np.sort(lp.cells), touching)
    assert total == s.ndofs


import pytest


@pytest.mark.parametrize("world,level", [(3, 1), (5, 2), (8, 2)])
def test_partition_merge_for_the_bench_world_sizes(world, level):
    """the 4- and 8-GPU runs of bench.py use the same host logic: every world size must reproduce the global matrix bit for bit
    (in-process, the oracle as the per-rank assembler)"""
    g = G.perturb_interior_nodes(G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), level))
    s = G.FESpace(G.H1P2(1, 3), g)
    blocks, cells = [], 0
    for r in range(world):
        lp = G.partition.partition(s, r, world)
        cells += lp.grid.ncells
        cp, rv, nz = _assemble_oracle(lp.grid, lp.space)
        blocks.append(G.partition.owned_block_to_global(lp, cp, rv, nz))
    colptr, rowval, nzval = G.partition.merge_owned_columns(s.ndofs, blocks)
    gcp, grv, gnz = _assemble_oracle(g, s)
    assert np.array_equal(colptr, gcp) and np.array_equal(rowval, grv) and np.array_equal(nzval, gnz)
    assert cells >= g.ncells          # halo cells are assembled more than once
