"""CPU-only check of the multi-rank host logic (block sharding + ordered gather of variable-length compressed
blocks), world_size 2 over gloo.  The codec itself is replaced by the oracle here (no GPU in this container); the
sharding/gather code under test is the same one bench.py runs over NCCL."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from bzip3_b200 import synth, sharding
    from tests import refs
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    data = synth.zipf_text(900_000, seed=7)
    bs = 1 << 17
    blocks = [data[i:i + bs] for i in range(0, len(data), bs)]
    mine = sharding.blocks_of_rank(len(blocks), rank, world)
    enc = {}
    for b in mine:
        e, r, err = refs.oracle_encode_block(blocks[b].tobytes(), bs)
        assert err == 0
        enc[b] = e
    gathered = sharding.gather_compressed(enc, len(blocks), rank, world, dist, device="cpu")
    if rank == 0:
        assert sorted(gathered) == list(range(len(blocks)))
        for b in range(len(blocks)):
            want = refs.oracle_encode_block(blocks[b].tobytes(), bs)[0]
            assert gathered[b] == want, b
        out = b"".join(refs.oracle_decode_block(gathered[b], len(blocks[b]), bs)[0] for b in range(len(blocks)))
        assert out == data.tobytes()
        print("GATHER_OK", len(blocks))
    dist.barrier()
    dist.destroy_process_group()
''') % ROOT


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = free_port()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GATHER_OK" in r.stdout
