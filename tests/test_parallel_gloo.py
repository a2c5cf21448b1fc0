"""Host logic of the frame-parallel vision stage on CPU: world_size 2, gloo (the N>1 path of bench.py / parallel.py)."""
import os
import socket

import pytest
import subprocess
import sys


def test_frame_shard_partition():
    from videollama2_b200.parallel import frame_shard, shard_sizes
    for F in (1, 2, 7, 16, 32):
        for W in (1, 2, 3, 4, 8):
            spans = [frame_shard(F, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = shard_sizes(F, W)
            assert sum(sizes) == F and max(sizes) - min(sizes) <= 1
    assert [frame_shard(16, r, 8) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]
    with pytest.raises(ValueError):
        frame_shard(4, 2, 2)


@pytest.mark.parametrize("F,world", [(16, 2), (7, 2), (1, 2), (7, 3), (16, 4)])
def test_all_gather_frames_gloo(F, world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gloo_worker.py")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, worker, str(F)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=150)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all(f"RANK{r} OK" in outs[r] for r in range(world))


@pytest.mark.parametrize("name,world", [("tiny", 2), ("tiny_qwen2", 1), ("mid", 2)])
def test_tp_partition_gloo(name, world):
    """Megatron-style partitioning of the decoder (tp_decoder.shard_plan / shard_state_dict): sharded fp32 mathematics with
    real all-reduce / all-gather over gloo reproduces the unsharded oracle forward."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_tp_gloo_worker.py")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, worker, name], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=200)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all(f"RANK{r} OK" in outs[r] for r in range(world))
