"""Full `cuobjdump -sass` listing of one kernel per family into profiles/sass/ (evidence that the tcgen05 / TMA /
multimem / peer-memory instructions are in the SAME kernel), plus a mnemonic census of every kernel in the library.
Runs on the CPU box (no GPU needed): python scripts/sass_listings.py"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "quintnet_b200", "lib", "libquintnet_b200.so")
out = os.path.join(root, "profiles", "sass")
os.makedirs(out, exist_ok=True)
FAMILIES = {   # file name -> mangled-name regex
    "gemm_bf16_cta_pair_plain": r"gemm_bf16_kernelILj256ELb0ELb0ELi0ELb1ELb0EEE",
    "gemm_bf16_cta_pair_FUSE_AG": r"gemm_bf16_kernelILj256ELb0ELb0ELi1ELb1ELb0EEE",
    "gemm_bf16_cta_pair_FUSE_RS": r"gemm_bf16_kernelILj256ELb0ELb0ELi2ELb1ELb0EEE",
    "gemm_bf16_wgrad_cta_pair": r"gemm_bf16_kernelILj256ELb1ELb1ELi0ELb1ELb0EEE",
    "attn_fwd": r"attn15attn_fwd_kernel", "attn_bwd": r"attn15attn_bwd_kernel",
    "small_attn_fwd": r"small_attn_fwd_kernel", "small_attn_bwd": r"small_attn_bwd_kernel",
    "ln_bwd_fused_bf16_dropout": r"ln_bwd_fused_kernelILi1ELb1EEE",
    "layernorm_fwd_bf16_dropout": r"layernorm_fwd_kernelI13__nv_bfloat16Li4ELb1EEE",
    "adamw_bf16param_f32grad": r"adamw_kernelI13__nv_bfloat16fEE",
    "ce_fwd_bwd": r"ce_fwd_bwd_kernel",
    "gemm_mxfp8_block_scaled": r"gemm_mxfp8_kernel", "quantize_mxfp8": r"quantize_mxfp8_kernel",
    "gemm_fp8_rowwise_cta_pair": r"gemm_bf16_kernelILj256ELb0ELb0ELi0ELb1ELb0ELb1EEE",
    "allreduce_two_shot_nvls_bf16": r"allreduce_kernelI13__nv_bfloat16Lb1EEE",
    "allreduce_two_shot_p2p_bf16": r"allreduce_kernelI13__nv_bfloat16Lb0EEE",
    "allreduce_oneshot_nvls_bf16": r"allreduce_oneshot_kernelI13__nv_bfloat16Lb1EEE",
    "reduce_scatter_sqnorm_nvls_f32": r"reduce_scatter_sqnorm_kernelIfLb1EEE",
    "adamw_allgather_nvls_bf16": r"adamw_allgather_kernelI13__nv_bfloat16Lb1EEE",
}
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n(?=\s+Function : )", txt)
census = []
KEYS = ["UTCHMMA", "UTCQMMA", "UTCCP", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "SYNCS", "LDGMC", "STG.E.128.STRONG.SYS",
        "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "LDG.E.STRONG.SYS", "STG.E.STRONG.SYS", "HMMA", "RED.E.ADD.F32"]
for b in blocks:
    m = re.match(r"\s+Function : (\S+)", b)
    if not m:
        continue
    name = m.group(1)
    insts = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+([^;]+);", b)
    cnt = {k: sum(1 for i in insts if re.search(r"(?<![A-Z])" + re.escape(k), i)) for k in KEYS}
    census.append((name, len(insts), cnt))
    for fname, rx in FAMILIES.items():
        if re.search(rx, name):
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            with open(os.path.join(out, fname + ".sass"), "w") as f:
                f.write(f"// {demangled}\n// cuobjdump -sass {os.path.relpath(so, root)}  (sm_100a)\n")
                keep = []
                for l in b.splitlines():
                    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?;)", l)      # address + instruction, encodings dropped
                    if mm:
                        keep.append(f"/*{mm.group(1)}*/  {mm.group(2)}")
                    elif "Function" in l or l.strip().startswith("."):
                        keep.append(l.strip())
                f.write("\n".join(keep) + "\n")
with open(os.path.join(out, "CENSUS.md"), "w") as f:
    f.write("# SASS mnemonic census of every kernel in libquintnet_b200.so (sm_100a)\n\n"
            "UTC*MMA = tcgen05.mma, UTCCP = tcgen05.cp (block-scale factors smem -> TMEM), LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = 1-D bulk async copy "
            "(the fused TP kernels' NVLink transport), LDGMC = multimem.ld_reduce (NVLS in-switch reduction); multimem.st has no mnemonic of its own — it is the "
            "`STG.E.128.STRONG.SYS` on the multicast-mapped address; "
            "`*.STRONG.SYS` / MEMBAR.*.SYS = system-scope peer flags.\n\n| kernel | SASS instrs | " + " | ".join(KEYS) + " |\n|---|---|" + "---|" * len(KEYS) + "\n")
    for name, n, cnt in sorted(census):
        dm = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dm = re.sub(r"\(.*", "", dm).replace("void ", "")
        f.write(f"| `{dm[:90]}` | {n} | " + " | ".join(str(cnt[k]) if cnt[k] else "" for k in KEYS) + " |\n")
print("wrote", len(os.listdir(out)), "files to", out)
