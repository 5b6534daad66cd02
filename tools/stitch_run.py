"""times hb_stitch_push under torchrun: torchrun --nproc-per-node N tools/stitch_run.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import huffman_gpu_b200 as hb
from huffman_gpu_b200 import sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = hb.workloads.get(sys.argv[1] if len(sys.argv) > 1 else "c5")
lo, hi = sharded.shard_bounds(wl.n_bytes // 4, world)[rank]
enc = hb.Encoder(local, (hi - lo) * 4)
comm = sharded.ShardComm(enc, rank, world)
d = torch.empty((hi - lo) * 4, dtype=torch.uint8, device="cuda")
enc.synth_fill(d, wl, first=lo * 4)
cw, cl, plan, _ = comm.plan_build(d)
loc = comm.local_buffer()
comm.encode_async(d, cw, cl, loc)
comm.encode_result()
comm.stitch_open(plan.total_bits // 32 + 2, 0)
comm.stitch_push(loc)
dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    comm.stitch_push(loc)
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / 5], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
mine = torch.tensor([(plan.shard_bits + 7) // 8 if rank else 0], dtype=torch.int64, device="cuda")
dist.all_reduce(mine)
if rank == 0:
    print("stitch %s over %d GPUs: %.3f ms, %.0f GB/s over NVLink (HB_STITCH_BLOCKS_PER_SM=%s)" % (
        wl.name, world, float(t[0]), int(mine[0]) / float(t[0]) / 1e6, os.environ.get("HB_STITCH_BLOCKS_PER_SM", "4")))
comm.stitch_close(); comm.close(); enc.close(); dist.destroy_process_group()
