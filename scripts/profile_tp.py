"""Per-kernel device-time of one optimizer step at tp=2 (rank 0), torch.profiler shares."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
sys.argv = [sys.argv[0], "--gpus", "2", "--mesh", "1,2,1"]
args = bench.parse_args()
args.global_batch = 128
rank, world, device = bench.dist_env(args)
mesh = bench.mesh_for(args)
from quintnet_b200 import GPT2Trainer, get_strategy, init_process_groups
from quintnet_b200.models.gpt2 import GPT2Config
cfg, M = bench.build_config(args, mesh, "ours")
pg = init_process_groups("cuda", cfg["mesh_dim"], cfg["mesh_name"])
model = get_strategy("3d", pg, cfg, checkpoint_path=None, is_staged=True).apply(None)
batches = bench.make_micro_batches(M, args.micro_batch, args.seq_len, args.vocab, 1, False, device)
tr = GPT2Trainer(model, bench.ListLoader(batches, args.micro_batch), None, cfg, pg, GPT2Config.from_dict(cfg["model_config"]))
for _ in range(3):
    tr.train_step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record(); t0 = time.perf_counter()
tr.train_step()
host_ms = (time.perf_counter() - t0) * 1e3          # host time to ENQUEUE the step (no sync inside = not host-bound if << step)
e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.device_time_total for e in evs)
    print(f"unprofiled step {step_ms:.1f} ms (host enqueue {host_ms:.1f} ms); sum of kernel time {tot/1e3:.1f} ms "
          f"for {M} micro-batches -> device busy {tot/1e3/step_ms*100:.0f}% (streams may overlap)")
    for e in sorted(evs, key=lambda e: -e.device_time_total)[:34]:
        print(f"{e.device_time_total/tot*100:5.1f}%  {e.device_time_total/1e3:9.2f} ms  x{e.count:5d}  {e.key[:100]}")
    # per-launch durations of the GEMM kernels (different layer shapes share a template instance)
    import collections
    per = collections.defaultdict(list)
    for e in prof.events():
        if "gemm_bf16_kernel" in e.name and e.device_time > 0:
            per[e.name[:60]].append(e.device_time)
    for k, v in per.items():
        v = sorted(v)
        q = [v[int(i * (len(v) - 1) / 10)] for i in range(11)]
        print(f"{k}  n={len(v)}  deciles(us)=" + " ".join(f"{x:.0f}" for x in q))
torch.distributed.barrier()
