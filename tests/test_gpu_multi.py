"""Multi-GPU (NCCL) parity: frame-chunk sharding with a one-sample halo + all-gather of records must give the
single-GPU result byte for byte; clip batches shard by clip.  Needs >= 2 GPUs (skipped otherwise); the same
logic runs on CPU/gloo in tests/test_host_logic.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r"""
import os, sys, json, numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from avd_b200 import dist as avd_dist, video
from avd_b200.engine import Engine
import avd_b200.synth as synth
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
eng = Engine(lrank)
# BASELINE configs[2] geometry at reduced length: 4K, 60 fps -> step 30
clip = synth.Clip(330, 2160, 3840, seed=2, fps=60.0)
frames = clip.sampled(30)                                   # 11 samples
meta = {{"fps": 60.0, "width": 3840, "height": 2160, "duration": 5.5}}
lo, hi, halo = avd_dist.local_range_with_halo(len(frames), rank, world)
res, full = avd_dist.analyze_frames_sharded(frames[lo:hi], frames[halo] if halo is not None else None, len(frames), meta,
                                            rank, world, eng.analyze_host, device=dev)
if rank == 0:
    single = eng.analyze_host(frames)
    assert full.tobytes() == single.tobytes(), "sharded records differ from the single-GPU run"
    assert res == video.analyze_frames(frames, meta, engine=eng)
    # clip batch, sharded by clip: every rank analyses its own clips, no exchange needed for the results
clips = [synth.Clip(60, 360, 640, seed=1000 + i, fps=30.0).sampled(15) for i in avd_dist.shard_clips(6, rank, world)]
outs = video.analyze_frames_batch(clips, [{{"fps": 30.0, "duration": 2.0}}] * len(clips), engine=eng)
for c, o in zip(clips, outs):
    assert o == video.analyze_frames(c, {{"fps": 30.0, "duration": 2.0}}, engine=eng)
dist.barrier()
if rank == 0:
    print("MULTI_OK", json.dumps(res["summary"]))
dist.destroy_process_group()
"""


def test_sharded_equals_single_gpu_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29547", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_OK" in out.stdout


_WORKER_ONE_GPU = r"""
import os, sys, json, numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from avd_b200 import dist as avd_dist, video
from avd_b200.engine import Engine
import avd_b200.synth as synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
eng = Engine(0)                                              # every rank computes on cuda:0; records travel over gloo
clip = synth.Clip(450, 1080, 1920, seed=2, fps=30.0)
frames = clip.sampled(15)                                    # 30 samples at BASELINE configs[1] geometry
meta = {{"fps": 30.0, "width": 1920, "height": 1080, "duration": 15.0}}
lo, hi, halo = avd_dist.local_range_with_halo(len(frames), rank, world)
res, full = avd_dist.analyze_frames_sharded(frames[lo:hi], frames[halo] if halo is not None else None, len(frames), meta,
                                            rank, world, eng.analyze_host, device=None)
single = eng.analyze_host(frames)
assert full.tobytes() == single.tobytes(), "sharded records differ from the single-GPU run"
assert res == video.analyze_frames(frames, meta, engine=eng)
dist.barrier()
if rank == 0:
    print("ONE_GPU_MULTI_OK")
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ranks_on_one_gpu_equal_single_run(tmp_path, world):
    """The sharded path end to end on a ONE-GPU box (where the NCCL test above has to skip): `world` processes, each
    analysing its chunk (+ one-sample halo) with the real kernels on cuda:0, records all-gathered over gloo; the
    assembled records must equal the single run byte for byte on every rank.  The kernels of the ranks are independent
    launches (nothing waits on another rank on the device)."""
    script = tmp_path / "w1.py"
    script.write_text(_WORKER_ONE_GPU.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29551 + world), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ONE_GPU_MULTI_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_clip_batch_equals_per_clip(synth):
    """Single GPU: one batched run over several clips == analysing each clip on its own."""
    from avd_b200 import video
    from avd_b200.engine import Engine
    eng = Engine(0)
    clips = [synth.Clip(75, 270, 480, seed=1000 + i, fps=30.0).sampled(15) for i in range(4)]
    metas = [{"fps": 30.0, "width": 480, "height": 270, "duration": 2.5}] * 4
    outs = video.analyze_frames_batch(clips, metas, engine=eng)
    for c, m, o in zip(clips, metas, outs):
        assert o == video.analyze_frames(c, m, engine=eng)
    assert video.analyze_frames_batch([], [], engine=eng) == []
    eng.close()
