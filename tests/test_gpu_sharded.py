"""Multi-GPU z-slab sharding (needs >= 2 GPUs; `gpurun --gpus 2`): NCCL frame broadcast, collective-
free integration, halo exchange + mesh gather -> identical to the single-GPU mesh."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret, layout, unit=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from bodyslam_b200 import ops
        from bodyslam_b200.sharding import ShardedTSDF
        from bodyslam_b200.tsdf import DenseTSDFVolume
        from util import canon_mesh, small_scene
        sc = small_scene("laparoscopy512", res=128, frames=6, with_color=False)
        sh = ShardedTSDF(sc["voxel_length"], sc["sdf_trunc"], 128, sc["origin"], color=False, device=dev, rank=rank, world_size=world, layout=layout, unit_activation=(unit is True), unit_arithmetic=(unit == 'arith'))
        F, H, W = sc["depth_u16"].shape
        if rank == 0:
            depth = ops.depth_from_u16(sc["depth_u16"], 1000.0, 3.0, dev)
            E = sc["E"]
        else:   # only rank 0 holds the frames: they arrive by NCCL broadcast
            depth = torch.zeros((F, H, W), dtype=torch.float32, device=dev)
            E = np.zeros_like(sc["E"])
        sh.integrate_batch(depth, None, sc["intrinsic"], E, broadcast_from=0)
        mesh = sh.extract_mesh()
        # the streamed replay (pinned host u16 on rank 0 -> H2D -> NCCL broadcast -> fused a4 + integrate,
        # pipelined over chunks) must build the same shard, and report the same per-frame update counts
        sh2 = ShardedTSDF(sc["voxel_length"], sc["sdf_trunc"], 128, sc["origin"], color=False, device=dev, rank=rank, world_size=world, layout=layout, unit_activation=(unit is True), unit_arithmetic=(unit == 'arith'))
        counts = torch.zeros(F, dtype=torch.int64, device=dev)
        src = torch.from_numpy(sc["depth_u16"]).pin_memory() if rank == 0 else None
        sh2.integrate_stream(src if rank == 0 else torch.empty(0), sc["intrinsic"], sc["E"] if rank == 0 else np.zeros_like(sc["E"]), src=0, chunk=2,
                             update_counts=counts)
        same = torch.equal(sh2.tsdf.export_dense()[0], sh.tsdf.export_dense()[0]) and torch.equal(sh2.tsdf.export_dense()[1], sh.tsdf.export_dense()[1])
        dist.all_reduce(counts)
        # sharded ingest: every rank feeds its share of each chunk from its own pinned host memory
        sh3 = ShardedTSDF(sc["voxel_length"], sc["sdf_trunc"], 128, sc["origin"], color=False, device=dev, rank=rank, world_size=world, layout=layout, unit_activation=(unit is True), unit_arithmetic=(unit == 'arith'))
        mine = ShardedTSDF.ingest_share(F, rank, world, 4)
        sh3.integrate_stream_sharded(torch.from_numpy(sc["depth_u16"][mine]).pin_memory(), sc["intrinsic"], sc["E"], chunk=4)
        same = same and torch.equal(sh3.tsdf.export_dense()[0], sh.tsdf.export_dense()[0]) and torch.equal(sh3.tsdf.export_dense()[1], sh.tsdf.export_dense()[1])
        # an odd frame count leaves a last chunk that does not divide evenly: per-piece broadcasts, device-resident shares
        sh4 = ShardedTSDF(sc["voxel_length"], sc["sdf_trunc"], 128, sc["origin"], color=False, device=dev, rank=rank, world_size=world, layout=layout, unit_activation=(unit is True), unit_arithmetic=(unit == 'arith'))
        mine = ShardedTSDF.ingest_share(5, rank, world, 4)
        sh4.integrate_stream_sharded(torch.from_numpy(sc["depth_u16"][mine]).to(dev), sc["intrinsic"], sc["E"][:5], chunk=4)
        sh5 = ShardedTSDF(sc["voxel_length"], sc["sdf_trunc"], 128, sc["origin"], color=False, device=dev, rank=rank, world_size=world, layout=layout, unit_activation=(unit is True), unit_arithmetic=(unit == 'arith'))
        sh5.integrate_batch(depth[:5].clone(), None, sc["intrinsic"], E[:5], broadcast_from=0)
        same = same and torch.equal(sh4.tsdf.export_dense()[0], sh5.tsdf.export_dense()[0]) and torch.equal(sh4.tsdf.export_dense()[1], sh5.tsdf.export_dense()[1])
        if rank == 0:
            ref = DenseTSDFVolume(sc["voxel_length"], sc["sdf_trunc"], 128, sc["origin"], color=False, device=dev, unit_activation=(unit is True), unit_arithmetic=(unit == 'arith'))
            ref.integrate_batch(depth, None, sc["intrinsic"], sc["E"])
            rm = ref.extract_triangle_mesh()
            a = canon_mesh(mesh.vertices.cpu().numpy(), mesh.vertex_keys.cpu().numpy(), mesh.triangles.cpu().numpy(), (128,) * 3)
            b = canon_mesh(rm.vertices.cpu().numpy(), rm.vertex_keys.cpu().numpy(), rm.triangles.cpu().numpy(), (128,) * 3)
            ok = np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])
            # the slab itself equals the slice of the full volume
            z0, z1 = sh.bounds[0]
            ok = ok and torch.equal(sh.contiguous_slab().export_dense()[1], ref.export_dense()[1][:, :, z0:z1])
            ok = ok and sh.layout == layout and same
            ok = ok and counts.cpu().tolist() == ref.count_updates(depth, sc["intrinsic"], sc["E"]).cpu().tolist()
            ret.put(("ok" if ok else "sharded mesh differs from single-GPU mesh", int(rm.triangles.shape[0])))
        else:
            assert mesh is None
            sh.contiguous_slab()      # collective: rank 0 calls it once more for the slab check
    except Exception as e:
        ret.put((f"rank {rank}: {type(e).__name__}: {e}", 0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("layout,unit", [("interleaved", False), ("contiguous", False), ("interleaved", True), ("interleaved", "arith")])
def test_sharded_mesh_equals_single_gpu_mesh(cuda, layout, unit):
    """unit=True: ScalableTSDFVolume unit activation on z-shards (the reference's TSDF() semantics, the default of
    both TSDF and ShardedTSDF on unit-aligned boxes) -- the sharded map / mesh equals the single-GPU drop-in's;
    unit="arith": dense rule with the reference's per-unit arithmetic on z-shards"""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret, layout, unit)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        msg, ntri = ret.get(timeout=240)
    except Exception:
        msg, ntri = "timeout waiting for the ranks", 0
    for p in procs:
        p.join(timeout=60 if msg == "ok" else 2)   # a failed rank leaves the others inside a collective: do not wait for them
        if p.is_alive():
            p.kill()
    assert msg == "ok", msg
    assert ntri > 5000
