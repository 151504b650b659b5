"""Per-kernel device-time shares of one optimizer step (torch.profiler; shares only — never a bench number)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

args = bench.parse_args()
args.global_batch = int(os.environ.get("QN_PROF_BATCH", 128))
mesh = (1, 1, 1)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29555")
rank, world, device = bench.dist_env(args)
from quintnet_b200 import GPT2Trainer, get_strategy, init_process_groups
from quintnet_b200.models.gpt2 import GPT2Config
cfg, M = bench.build_config(args, mesh, "ours")
pg = init_process_groups("cuda", cfg["mesh_dim"], cfg["mesh_name"])
model = get_strategy("3d", pg, cfg, checkpoint_path=None, is_staged=True).apply(None)
batches = bench.make_micro_batches(M, args.micro_batch, args.seq_len, args.vocab, 1, False, device)
tr = GPT2Trainer(model, bench.ListLoader(batches, args.micro_batch), None, cfg, pg, GPT2Config.from_dict(cfg["model_config"]))
for _ in range(2):
    tr.train_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_step()
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in evs)
rows = sorted(evs, key=lambda e: -e.device_time_total)[:40]
out = [{"kernel": e.key[:110], "calls": e.count, "total_ms": e.device_time_total / 1e3, "share": e.device_time_total / tot} for e in rows]
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"total_device_ms": tot / 1e3, "micro_batches": M, "top": out}, open("gpurun_out/profile_step.json", "w"), indent=1)
print(f"total device ms {tot/1e3:.1f} for {M} micro-batches")
for r in out:
    print(f"{r['share']*100:5.1f}%  {r['total_ms']:9.2f} ms  x{r['calls']:5d}  {r['kernel']}")
