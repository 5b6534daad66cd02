"""debug aid: push-stitched vs direct-encoded stream, torchrun --nproc-per-node N tools/direct_debug.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import huffman_gpu_b200 as hb
from huffman_gpu_b200 import sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = hb.workloads.get(sys.argv[1] if len(sys.argv) > 1 else "c4")
lo, hi = sharded.shard_bounds(wl.n_bytes // 4, world)[rank]
enc = hb.Encoder(local, (hi - lo) * 4)
comm = sharded.ShardComm(enc, rank, world)
d = torch.empty((hi - lo) * 4, dtype=torch.uint8, device="cuda")
enc.synth_fill(d, wl, first=lo * 4)
cw, cl, plan, _ = comm.plan_build(d)
starts, bits = comm.offsets()
loc = comm.local_buffer()
comm.encode_async(d, cw, cl, loc)
comm.encode_result()
cap = plan.total_bits // 32 + 2
comm.stitch_open(cap, 0)
comm.stitch_push(loc)
torch.cuda.synchronize(); dist.barrier()
n = plan.total_bits // 32 + 1
if rank == 0:
    ref = comm.stitched_view(cap).clone()
    comm.stitched_view(cap).fill_(0x5A5A5A5A)
torch.cuda.synchronize(); dist.barrier()
comm.encode_direct_async(d, cw, cl)
print("rank", rank, "bits", comm.encode_result(), plan.shard_bits, "start", int(starts[rank]), "phase", int(starts[rank]) % 32, "end&31", (int(starts[rank]) + int(bits[rank])) % 32, flush=True)
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    got = comm.stitched_view(cap)
    bad = torch.nonzero(got[:n] != ref[:n]).flatten()
    print("mismatching words:", bad.numel(), "of", n)
    if bad.numel():
        b = bad.cpu().numpy()
        print("first", b[:10], "last", b[-10:])
        seams = [int(s) // 32 for s in starts]
        print("seam words", seams)
        for i in b[:6]:
            print(int(i), "got %08x want %08x" % (int(got[i].item()) & 0xFFFFFFFF, int(ref[i].item()) & 0xFFFFFFFF))
        runs = np.split(b, np.nonzero(np.diff(b) != 1)[0] + 1)
        print("runs:", len(runs), [(int(r[0]), len(r)) for r in runs[:10]])
comm.stitch_close(); comm.close(); enc.close(); dist.destroy_process_group()
