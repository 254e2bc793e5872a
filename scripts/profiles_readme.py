"""Regenerates the measured tables of profiles/README.md (between the ROUND2 markers) from the JSON artefacts under
profiles/: bench lines, the CUDA-event kernel table, the ncu full-metric summary.  python scripts/profiles_readme.py"""
import json, os
R = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
J = lambda f: json.load(open(os.path.join(R, f)))


def line(f):
    return json.loads(open(os.path.join(R, f)).read().strip().split("\n")[-1])


out = []
b1 = line("r2_bench_1gpu.json")
rows = [("1 GPU, 224x224, batch 256 (BASELINE configs[1])", "r2_bench_1gpu.json"), ("2 GPUs (NCCL, weak scaling)", "r2_bench_2gpu.json"),
        ("4 GPUs (NCCL, weak scaling)", "r2_bench_4gpu.json"),
        ("8 GPUs (NCCL, weak scaling; measured one kernel-policy change earlier, at 14.9 ms/step on one GPU)", "r2_bench_8gpu.json"),
        ("1 GPU, 112x112, batch 512 (configs[3])", "r2_bench_cfg3_112_b512.json"), ("1 GPU, 192x256, batch 256 (configs[4])", "r2_bench_cfg4_192x256.json"),
        ("1 GPU, 256x192, batch 256 (configs[4])", "r2_bench_cfg4_256x192.json"),
        ("1 GPU, 224x224, 2048 images per step (configs[2] strong-scaling shard)", "r2_bench_cfg2_strong_2048_1gpu.json")]
out.append("| run | images/s (`value`) | ms/step | e2e images/s (host batches, H2D inside the timed region) | SM clock under load | file |")
out.append("|---|---|---|---|---|---|")
for name, f in rows:
    if not os.path.exists(os.path.join(R, f)): continue
    d = line(f)
    e = d.get("e2e") or {}
    c = d.get("clocks") or {}
    out.append(f"| {name} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {(e.get('value') or 0):.0f} | {(c.get('sm_mhz') or 0):.0f} MHz, reasons {c.get('reasons', [])} | `{f}` |")
ref = line("r2_bench_reference_arm.json")
ge = b1.get("gpu_eager_baseline") or {}
cb = b1.get("cpu_baseline") or {}
out.append("")
out.append(f"* Reference arm (`bench.py --impl reference`, the reference's own torch CPU kernels through the oracle port, all host cores): "
           f"**{ref['value']:.1f} images/s**; `cpu_baseline` inside the 1-GPU line: {cb.get('value', 0):.1f} images/s on {cb.get('cores', '?')} cores ({cb.get('sample', '')}).")
out.append(f"* Stock PyTorch on the same B200 (`gpu_eager_baseline` inside the 1-GPU line; {ge.get('what', '')}): **{ge.get('value', 0):.0f} images/s**.")
rf = b1["roofline"]
out.append(f"* `roofline` of the 1-GPU line: kernel `{rf.get('kernel')}`, bound {rf['bound']}, achieved {rf['achieved']:.0f} {rf['unit']} of {rf['peak']:.1f} "
           f"(measured peak, MEASURED_PEAKS.json) = **{rf['frac']:.3f}**, traffic {rf.get('traffic')}; `gpu_launches` {b1.get('gpu_launches')} in {b1['steps']} timed steps.")
out.append("")
kt = J("r2_kernel_table.json")
out.append(f"### Families (`r2_kernel_table.json`: CUDA events around every launch, single stream, sum {kt['ms_per_step_sum']:.2f} ms; the graph replay overlaps the weight-gradient stream and takes {b1['ms_per_step']:.2f} ms)")
out.append("")
out.append("| family | ms/step | share | section-8d bytes GB/step | achieved GB/s | of measured HBM peak | GFLOP/step | TFLOP/s | DRAM GB/step (ncu) |")
out.append("|---|---|---|---|---|---|---|---|---|")
for k, v in kt["families"].items():
    out.append(f"| {k} | {v['ms_per_step']:.2f} | {100 * v['share']:.1f} % | {v.get('algorithmic_GB', 0):.2f} | {v.get('achieved_GBps', 0):.0f} | {v.get('frac_of_hbm_peak', 0):.2f} | "
               f"{v.get('algorithmic_GFLOP', 0):.0f} | {v.get('achieved_TFLOPs', 0):.1f} | {v.get('dram_GB_ncu', 0)} |")
out.append("")
out.append("### Kernel classes (same file)")
out.append("")
out.append("| kernel class | family | launches/step | ms/step | share | kernel traffic GB/step | achieved GB/s | of measured HBM peak |")
out.append("|---|---|---|---|---|---|---|---|")
for k in kt["kernels"]:
    if k["ms_per_step"] < 0.02: continue
    out.append(f"| `{k['kernel']}` | {k['family']} | {k['launches_per_step']} | {k['ms_per_step']:.3f} | {100 * k['share']:.1f} % | {k['kernel_traffic_GB_per_step']:.2f} | "
               f"{k['achieved_GBps']:.0f} | {k['frac_of_hbm_peak']:.2f} |")
out.append("")
out.append("### Largest layers (same file; one row per (kernel class, layer))")
out.append("")
out.append("| kernel class | layer | launches/step | us/step | achieved GB/s (section-8d bytes) |")
out.append("|---|---|---|---|---|")
for r in sorted(kt["layers"], key=lambda r: -r["ms_per_step"])[:28]:
    out.append(f"| `{r['kernel']}` | {r['layer']} | {r['launches_per_step']} | {1000 * r['ms_per_step']:.0f} | {r['achieved_GBps']:.0f} |")
out.append("")
nc = J("r2_ncu_gemm_dw_summary.json")
out.append(f"### ncu `--set full` of {len(nc)} consecutive GEMM / depthwise / fused launches of a timed step (`r2_ncu_gemm_dw_summary.json`, one row per distinct kernel)")
out.append("")
out.append("| kernel | us | tensor pipe active % | issue active % | warps active % | regs | smem KB | DRAM MB (read + write) | top stalls per issue |")
out.append("|---|---|---|---|---|---|---|---|---|")
seen = set()
for r in nc:
    k = r["kernel"].replace("void ", "").split("(")[0]
    if k in seen: continue
    seen.add(k)
    st = ", ".join(f"{a} {b:.1f}" for a, b in list(r["stalls_per_issue"].items())[:3])
    out.append(f"| `{k}` | {r['dur_us']:.0f} | {r['tensor_pipe_active_pct']:.1f} | {r['issue_active_pct']:.1f} | {r['warps_active_pct']:.1f} | {r['regs']:.0f} | "
               f"{r['smem_dyn_kb']:.0f} | {r['dram_read_mb'] + r['dram_write_mb']:.0f} | {st} |")
nb = J("r2_ncu_gemm_dw_summary_backward.json")
out.append("")
out.append(f"### The same ncu pass over the BACKWARD half of a step ({len(nb)} launches, captured at the 15.09 ms build: same kernel code; since then every depthwise fused backward runs `dws_bwd_k`) -- `r2_ncu_gemm_dw_summary_backward.json`, kernels not in the table above")
out.append("")
out.append("| kernel | us | tensor pipe active % | issue active % | warps active % | regs | smem KB | DRAM MB (read + write) | top stalls per issue |")
out.append("|---|---|---|---|---|---|---|---|---|")
for r in nb:
    k = r["kernel"].replace("void ", "").split("(")[0]
    if k in seen: continue
    seen.add(k)
    st = ", ".join(f"{a} {b:.1f}" for a, b in list(r["stalls_per_issue"].items())[:3])
    out.append(f"| `{k}` | {r['dur_us']:.0f} | {r['tensor_pipe_active_pct']:.1f} | {r['issue_active_pct']:.1f} | {r['warps_active_pct']:.1f} | {r['regs']:.0f} | "
               f"{r['smem_dyn_kb']:.0f} | {r['dram_read_mb'] + r['dram_write_mb']:.0f} | {st} |")
frag = "\n".join(out)
p = os.path.join(R, "README.md")
s = open(p).read()
a, b = "<!-- ROUND2-TABLES-BEGIN -->", "<!-- ROUND2-TABLES-END -->"
if a in s:
    s = s[:s.index(a) + len(a)] + "\n" + frag + "\n" + s[s.index(b):]
    open(p, "w").write(s)
    print("profiles/README.md updated,", len(out), "lines")
else:
    print(frag)
