"""Data-parallel host logic with world_size 2 over gloo on CPU: parameter broadcast,
gradient all-reduce on the flat buffers, 1/world folded into the optimizer, disjoint shards."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, workdir, out):
    import sys
    for p in (ROOT, os.path.join(ROOT, "vae-gam_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import vae_reg_GP
    from vaegam import dp, synthetic as syn
    r, w, _ = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    tr, te, glm = (os.path.join(workdir, f) for f in ("train.csv", "test.csv", "glm_maps.csv"))
    torch.manual_seed(100 + rank)              # ranks start from DIFFERENT parameters
    model = vae_reg_GP.VAE(save_dir=os.path.join(workdir, f"r{rank}"), glm_maps=glm, csv_files=[tr, te],
                           device_name="cpu")
    red = dp.GradientAllReduce(model._flat, model.optimizer, overlap=False)
    red.broadcast_parameters(0)
    p_sum = float(model._flat.flat32.double().sum())
    model._flat.grad32.fill_(float(rank + 1))
    model._flat.grad64.fill_(float(10 * (rank + 1)))
    red()
    shard = dp.shard_indices(1372, rank, world, epoch=3, seed=5)
    out[rank] = {"p_sum": p_sum, "g32": float(model._flat.grad32[0]), "g64": float(model._flat.grad64[-1]),
                 "scale": model.optimizer.grad_scale, "shard": shard.tolist(),
                 "fc1_is_view": model.fc1.weight.data_ptr() == model._flat.flat32.data_ptr() + 4 * model._flat.slices["fc1.weight"][1]}
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_gloo():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "vae-gam_b200"))
    from vaegam import synthetic as syn
    workdir = tempfile.mkdtemp(prefix="dp_")
    syn.write_experiment(workdir, n_subjects=2, config="checker")
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, workdir, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert a["p_sum"] == b["p_sum"]                       # rank-0 broadcast made parameters identical
    assert a["g32"] == b["g32"] == 3.0 and a["g64"] == b["g64"] == 30.0
    assert a["scale"] == b["scale"] == 0.5                # mean over ranks is applied inside the fused Adam
    assert a["fc1_is_view"] and b["fc1_is_view"]
    sa, sb = set(a["shard"]), set(b["shard"])
    assert len(sa) == len(sb) == 686 and not (sa & sb)


def test_shards_cover_and_reshuffle():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "vae-gam_b200"))
    from vaegam import dp
    parts = [dp.shard_indices(100, r, 4, epoch=0) for r in range(4)]
    assert sorted(torch.cat(parts).tolist()) == list(range(100))
    assert dp.shard_indices(100, 0, 4, epoch=1).tolist() != parts[0].tolist()
    assert dp.shard_indices(10, 1, 2, shuffle=False).tolist() == [5, 6, 7, 8, 9]
