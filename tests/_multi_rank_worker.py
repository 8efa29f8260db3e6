"""One rank of the one-process-per-GPU shape of lp_multi (lp_multi_create_rank), run as a
subprocess by tests/test_gpu_multi.py:  python _multi_rank_worker.py RANK WORLD DIR SPP [auto|nccl|peer]
Rank 0 writes the NCCL id to DIR/id.bin (the "own means" by which the id travels), every rank
renders its share of the cornell box, rank 0 saves the reduced frame to DIR/out.npz."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import loupiote_b200 as lb  # noqa: E402
from loupiote_b200 import scenes  # noqa: E402


def main():
    rank, world, out_dir, spp = int(sys.argv[1]), int(sys.argv[2]), Path(sys.argv[3]), int(sys.argv[4])
    mode = sys.argv[5] if len(sys.argv) > 5 else "auto"
    id_path = out_dir / "id.bin"
    if rank == 0:
        uid = lb.MultiRenderer.unique_id()
        tmp = out_dir / "id.tmp"
        tmp.write_bytes(uid)
        tmp.rename(id_path)
    else:
        t0 = time.time()
        while not id_path.exists():
            if time.time() - t0 > 60:
                raise SystemExit("no NCCL id from rank 0")
            time.sleep(0.01)
        uid = id_path.read_bytes()
    m = lb.MultiRenderer.create_rank(rank, uid, world, rank)
    assert (m.world, m.rank, m.local_devices) == (world, rank, 1)
    c = scenes.cornell_box()
    m.set_scene(c["scene"])
    m.resize((64, 48))
    m.resize((96, 64))  # a second resize: the peers' IPC mappings are dropped and re-made
    if mode == "nccl":
        m.set_reduce_mode(lb.ReduceMode.NCCL)
    elif mode == "peer":
        if not m.peer_access:
            raise SystemExit(77)  # no peer access between the two GPUs: the test skips
        m.set_reduce_mode(lb.ReduceMode.PEER)
    m.set_config(max_bounces=4, seed=7, spp_per_call=spp)
    for _ in range(2):  # two batches back to back: the second overwrites, no sync between
        m.render(c["view"])
        m.reduce()
    if rank == 0:
        acc = m.read_accum_sum()
        np.savez(out_dir / "out.npz", accum=acc, pixels=m.read_pixels(), peer=np.array(m.peer_access),
                 counters=np.array([m.ray_counters()[k] for k in ("primary", "bounce", "shadow")]))
    else:
        m.synchronize()
    m.close()


if __name__ == "__main__":
    main()
