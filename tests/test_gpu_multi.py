"""Two-GPU checks of the partitioned path (skipped on a single-GPU box): the peer-memory exchange against NCCL, and the
partitioned coarse solve + element solves against the single-GPU result, by stream launches and replayed as a CUDA graph.
One process per GPU, NCCL for the plumbing."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, E, out_dir):
    import torch.distributed as dist
    from hybrid_fem_lssvr_b200 import batch, dist as hdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        px = hdist.PeerExchange(device=dev)
        # 1. all-gather: peer memory against NCCL, several epochs and both widths
        g = torch.Generator(device='cpu').manual_seed(rank)
        for it in range(4):
            for ch, W in ((px.CHANNEL_INTERFACE, 4), (px.CHANNEL_ERROR, 3)):
                src = torch.randn(W, generator=g, dtype=torch.float64).to(dev)
                a = px.all_gather(src, ch)
                b = torch.empty((world, W), dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(b.reshape(-1), src)
                assert torch.equal(a, b), (rank, it, ch)
        # 2. partitioned step: same bc2 / fine grid / error through both transports
        nodes = hdist.local_nodes_linspace(-1.0, 1.0, E * world, world, rank, device=dev)
        res = {}
        for name, ex in (('peer', px), ('nccl', None)):
            y, bc2 = hdist.fem_p1_solve_distributed(nodes, exchange=ex)
            err = batch.new_error_accumulator(dev)
            _, fine, _ = batch.lssvr_primal_batch(nodes, y, 9, 1e4, N=12, F=32, bc2=bc2, want_coef=False, want_fine=True, err3=err)
            l2, mx, failed = hdist.finish_gathered_error(hdist.gather_error(err, exchange=ex))
            res[name] = (bc2.clone(), fine, l2, mx, failed)
        assert torch.equal(res['peer'][0], res['nccl'][0]) and torch.equal(res['peer'][1], res['nccl'][1])
        # the L2 accumulator is a floating-point atomicAdd: the two runs may differ in the last bits
        assert abs(res['peer'][2] - res['nccl'][2]) <= 1e-9 * res['nccl'][2] and res['peer'][3:] == res['nccl'][3:]
        assert not px.timed_out()
        # 3. the same step captured once and replayed as a CUDA graph (the exchange keeps its epoch on the device, so the
        # captured launch is valid for every later step): same bits as the stream launches, replay after replay
        u_g = torch.empty_like(nodes)
        fine_g = torch.empty_like(res['peer'][1])
        err_g = batch.new_error_accumulator(dev)
        def step():
            _, b = hdist.fem_p1_solve_distributed(nodes, exchange=px, out=u_g)
            batch.lssvr_primal_batch(nodes, u_g, 9, 1e4, N=12, F=32, bc2=b, want_coef=False, want_fine=True, fine_out=fine_g,
                                     err3=err_g)
            return b
        step()
        torch.cuda.synchronize()
        dist.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            bc2_g = step()
        for _ in range(3):
            fine_g.fill_(float('nan'))
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(bc2_g, res['peer'][0]), (rank, bc2_g.tolist(), res['peer'][0].tolist())
            assert torch.equal(fine_g, res['peer'][1]), (rank, float((fine_g - res['peer'][1]).abs().max()))
        assert not px.timed_out()
        np.save(os.path.join(out_dir, 'fine_%d.npy' % rank), res['peer'][1].cpu().numpy())
        if rank == 0:
            np.save(os.path.join(out_dir, 'err.npy'), np.array(res['peer'][2:4]))
        px.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_partitioned_step_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from hybrid_fem_lssvr_b200 import batch
    world, E = 2, 50000
    mp.spawn(_worker, args=(world, _free_port(), E, str(tmp_path)), nprocs=world, join=True)
    fine = np.concatenate([np.load(tmp_path / ('fine_%d.npy' % r)) for r in range(world)])
    nodes = batch.mesh_linspace(-1.0, 1.0, E * world + 1)
    u = batch.fem_p1_solve(nodes, coarse_solver='assembled_exact')
    err = batch.new_error_accumulator()
    _, ref, _ = batch.lssvr_primal_batch(nodes, u, 9, 1e4, N=12, F=32, want_coef=False, want_fine=True, err3=err)
    assert np.max(np.abs(fine - ref.cpu().numpy())) <= 1e-12
    l2, mx = batch.finish_error(err)
    e = np.load(tmp_path / 'err.npy')
    assert abs(e[0] - l2) <= 1e-3 * l2 + 1e-14 and abs(e[1] - mx) <= 1e-3 * mx + 1e-14
