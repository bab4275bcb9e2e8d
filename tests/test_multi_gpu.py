"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`): the
collective predict of the C ABI (reference rows sharded over the ranks, reads hashed 1/world per rank, NCCL exchange of
the query lists and of the local top-N) == the oracle, for several `top`, chunkings and ranking modes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_predict_matches_oracle(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(HERE, "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") >= 3 and "MISMATCH" not in r.stdout, r.stdout[-2000:]


def test_cli_two_ranks_equal_one(tmp_path):
    """The C++ `sketchy` binary as two processes (one per GPU, NCCL id through a file): `sketch` partitions the input files
    over the ranks and writes the same .msh; `predict` (streaming and read-set mode) shards the reference rows and prints
    the same rows as a single process."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    sys.path.insert(0, os.path.dirname(HERE))
    from sketchy_b200 import synth
    from sketchy_b200 import build as skb_build
    exe = skb_build.CLI
    base = [synth.random_genome(40_000, 600 + i) for i in range(3)]
    paths = []
    for g in range(9):
        p = tmp_path / f"g{g}.fa"
        p.write_bytes(b">g%d\n" % g + synth.mutate(base[g % 3], 0.002, 700 + g).tobytes() + b"\n")
        paths.append(str(p))
    geno = tmp_path / "g.tsv"
    geno.write_text("id\tmlst\n" + "".join(f"g{g}.fa\tST{g % 3}\n" for g in range(9)))
    blob, roff, _ = synth.sample_reads(base, 333, 1200, 41)
    fq = tmp_path / "r.fq"
    with open(fq, "w") as f:
        for i in range(333):
            r = blob[int(roff[i]):int(roff[i + 1])].tobytes().decode()
            f.write(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n")

    def run(world, *args):
        comm = str(tmp_path / f"comm_{world}_{len(os.listdir(tmp_path))}")
        procs = []
        for r in range(world):
            env = dict(os.environ, SKETCHY_B200_RANK=str(r), SKETCHY_B200_WORLD=str(world), SKETCHY_B200_DEVICE=str(r),
                       SKETCHY_B200_COMM_FILE=comm)
            procs.append(subprocess.Popen([exe, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env))
        outs = [p.communicate(timeout=600) for p in procs]
        assert all(p.returncode == 0 for p in procs), [o[1][-500:] for o in outs]
        assert all(o[0] == b"" for o in outs[1:])          # only rank 0 prints
        return outs[0][0]

    one, two = tmp_path / "one.msh", tmp_path / "two.msh"
    run(1, "sketch", "-i", *paths, "-o", str(one), "-s", "300", "-k", "16")
    run(2, "sketch", "-i", *paths, "-o", str(two), "-s", "300", "-k", "16")
    assert one.read_bytes() == two.read_bytes()
    for extra in (["-s", "-t", "3"], ["-s", "-t", "3", "-c"], ["-t", "4"]):
        a = run(1, "predict", "-i", str(fq), "-r", str(one), "-g", str(geno), *extra)
        b = run(2, "predict", "-i", str(fq), "-r", str(one), "-g", str(geno), *extra)
        assert a == b and a.count(b"\n") >= 1, (extra, a[:300], b[:300])
