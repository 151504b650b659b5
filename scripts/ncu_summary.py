"""Turn an `ncu --set full` report into the committed evidence: one markdown row per profiled launch with duration,
tensor-pipe activity, DRAM / L2->SM bytes against the algorithmic bytes, registers, and the achieved fraction of the
MEASURED peaks (MEASURED_PEAKS.json).  usage: python scripts/ncu_summary.py gpurun_out/r2_prof.ncu-rep profiles/r2_ncu_summary.md"""
import csv, io, json, os, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_fallback": True}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
nvtx_col = next(i for i, h in enumerate(hdr) if h.startswith("thread Domain:Push/Pop_Range"))
# algorithmic work of each target (name -> (flops, bytes)); see scripts/ncu_r2_targets.py
M, E = 16384, 1024
ALG = {
    "gemm_fwd_16384x4096x1024": (2.0 * M * 4096 * E, (M * E + 4096 * E + M * 4096) * 2),
    "gemm_bias_gelu_16384x4096x1024": (2.0 * M * 4096 * E, (M * E + 4096 * E + 2 * M * 4096) * 2),
    "gemm_dgrad_16384x1024x4096": (2.0 * M * 4096 * E, (M * 4096 + 4096 * E + M * E) * 2),
    "gemm_wgrad_accum_f32_4096x1024x16384": (2.0 * M * 4096 * E, (M * 4096 + M * E) * 2 + 4096 * E * 4),
    "attn_fwd_B32H16T512": (4.0 * 32 * 16 * 512 * 512 * 64 / 2, 32 * 512 * 16 * 64 * 2 * 4),
    "attn_bwd_B32H16T512": (4.0 * 32 * 16 * 512 * 512 * 64 / 2 * 2.5, 32 * 512 * 16 * 64 * 2 * 8),
    "dropout_add_layernorm_fwd_16384x1024": (0, 4 * M * E * 2),
    "dropout_add_layernorm_bwd_16384x1024": (0, 5 * M * E * 2),
    "adamw_64Mi_bf16param_f32grad": (0, (64 << 20) * 32),
    "ce_fwd_bwd_8192x50257": (0, 3 * 8192 * 50257 * 2),
    "fused_ag_gemm_pump_16384x1536x1024": (2.0 * M * 1536 * E, 0),
    "fused_ag_gemm_lsu_16384x1536x1024": (2.0 * M * 1536 * E, 0),
    "fused_gemm_rs_16384x1024x2048": (2.0 * M * E * 2048, 0),
    "plain_gemm_16384x1536x1024": (2.0 * M * 1536 * E, 0),
    "plain_gemm_16384x1024x2048": (2.0 * M * E * 2048, 0),
}
def g(r, name, default=""):
    return r[col[name]] if name in col else default
lines = ["| target (NVTX range) | kernel | duration µs | SM clock GHz | tensor pipe active % | achieved TFLOP/s (× of measured sustained / burst) | "
         "DRAM rd+wr MB (algorithmic MB) | achieved GB/s (× of measured copy) | L2→SM MB | regs/thr | grid |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows[2:]:
    tag = r[nvtx_col]
    name = tag.split(":")[1] if ":" in tag else ""
    kern = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("qn::", "")[:60]
    if "elementwise" in kern or "distribution" in kern or "finalize" in kern or "Fill" in kern:
        continue
    dur = float(g(r, "gpu__time_duration.sum", "0") or 0)            # us
    ghz = float(g(r, "sm__cycles_elapsed.avg.per_second", "0") or 0)
    tens = g(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "") or \
        g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "")
    dr = float(g(r, "dram__bytes_read.sum", "0") or 0); dw = float(g(r, "dram__bytes_write.sum", "0") or 0)
    unit_r = rows[1][col["dram__bytes_read.sum"]]; unit_w = rows[1][col["dram__bytes_write.sum"]]
    mul = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}
    dmb = dr * mul.get(unit_r, 1.0) + dw * mul.get(unit_w, 1.0)
    x2l = g(r, "l1tex__m_xbar2l1tex_read_bytes.sum", "")
    x2l_mb = float(x2l) * mul.get(rows[1][col["l1tex__m_xbar2l1tex_read_bytes.sum"]], 1.0) if x2l else 0.0
    fl, by = ALG.get(name, (0, 0))
    tf = fl / dur / 1e6 if dur and fl else 0.0
    gbs = by / dur / 1e3 if dur and by else 0.0
    tfs = f"{tf:.0f} ({tf / peaks['bf16_tflops_sustained']:.2f} / {tf / peaks['bf16_tflops']:.2f})" if tf else ""
    gbss = f"{gbs:.0f} ({gbs / peaks['hbm_gbs']:.2f})" if gbs else ""
    lines.append(f"| {name} | `{kern}` | {dur:.1f} | {ghz:.2f} | {tens[:5]} | {tfs} | {dmb:.0f} ({by / 1e6:.0f}) | {gbss} | "
                 f"{x2l_mb:.0f} | {g(r, 'launch__registers_per_thread')} | {g(r, 'launch__grid_size')} |")
hdr_txt = (f"# ncu --set full summary ({os.path.basename(rep)})\n\nDenominators: MEASURED_PEAKS.json "
           f"{'(FALLBACK values)' if peaks.get('_fallback') else ''}: bf16 {peaks['bf16_tflops_sustained']} TFLOP/s sustained / "
           f"{peaks['bf16_tflops']} burst, HBM copy {peaks['hbm_gbs']} GB/s.  Durations are ncu's `gpu__time_duration` of ONE isolated, "
           "cache-flushed replay at the clock ncu saw (column 4) — shares and ratios, not bench numbers; CUDA-event timings of the same "
           "launches are in `r2_kernel_times.json`.\n\n")
open(out, "w").write(hdr_txt + "\n".join(lines) + "\n")
print(hdr_txt + "\n".join(lines))
