"""GPU, 2 ranks over NCCL (needs >= 2 GPUs: gpurun --gpus 2): shot- and model-sharded gradients equal the
single-GPU gradient of the same batch."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT, Golden, rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from red_diffeq_b200 import ShardedFWIForward, s_normalize_none, v_denormalize
    g = Golden("tiny_half_receivers")
    dev = torch.device("cuda", rank)
    op = ShardedFWIForward(g.fresh_ctx(), dev, mode=mode, sample_temporal=g.sample_temporal, sample_spatial=g.sample_spatial,
                           normalize=True, v_denorm_func=v_denormalize, s_norm_func=s_normalize_none)
    v = torch.tensor(g.v, device=dev, requires_grad=True)
    seis = op(v)
    cot = torch.tensor(g.cotangent((g.v.shape[0], 4, 70, 10)), device=dev)
    (seis * op.local_slice(cot)).sum().backward()
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"grad{rank}.npy"), v.grad.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["models", "shots"])
def test_two_gpu_gradient_matches_reference(tmp_path, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = Golden("tiny_half_receivers")
    mp.spawn(_worker, args=(2, 29711 + (mode == "shots"), mode, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "grad0.npy"), np.load(tmp_path / "grad1.npy")
    assert np.array_equal(g0, g1)
    assert rel_l2(g0, g.grad_f32) <= 1e-4
