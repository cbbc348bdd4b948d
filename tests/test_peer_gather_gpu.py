"""Two-GPU test of the peer-to-peer all-gather of the packed results (needs >= 2 visible GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from specdec_b200 import dist as sdd
total, width = 64, 6
pg = sdd.PeerGather(total, width, slots=4)
lo, hi = sdd.shard_range(total, rank, world)
for step in range(11):  # wraps the ring of 4 slots twice
    g = torch.Generator().manual_seed(100 + step)
    full = torch.randint(-1, 1000, (total, width), generator=g, dtype=torch.int32).cuda()
    pg.publish(full[lo:hi].contiguous(), step)
    got = pg.gathered(step)
    torch.cuda.synchronize()
    assert torch.equal(got, full), (rank, step)
    ref = sdd.all_gather_packed(full[lo:hi].contiguous(), total)
    assert torch.equal(ref, full)
    if step >= 1:  # the fused form: publish step s and read step s - 1 in one launch (here: re-publish the same rows)
        prev = pg.publish(full[lo:hi].contiguous(), step, step - 1)
        torch.cuda.synchronize()
        gp = torch.Generator().manual_seed(100 + step - 1)
        assert torch.equal(prev, torch.randint(-1, 1000, (total, width), generator=gp, dtype=torch.int32).cuda())
    dist.barrier()  # (a reader is done with the slot before anyone can wrap around to it)
assert int(pg.status[0]) == 0
dist.barrier()
# overlap mode: the publish kernel rides a side stream behind an event on the caller's stream
po = sdd.PeerGather(total, width, slots=4, overlap=True)
for step in range(6):
    g = torch.Generator().manual_seed(500 + step)
    full = torch.randint(-1, 1000, (total, width), generator=g, dtype=torch.int32).cuda()
    po.publish(full[lo:hi].contiguous(), step, step)
    po.sync_reader()
    torch.cuda.synchronize()
    assert torch.equal(po.buf[step % 4], full), (rank, step)
    dist.barrier()
assert int(po.status[0]) == 0
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_gather_equals_nccl_all_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29683")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0 and "ok" in out, out
