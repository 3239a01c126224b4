#!/usr/bin/env python
"""Tables of profiles/r2_summary.md, regenerated from the JSON records in this directory (nothing typed by hand):
    python profiles/make_r2_summary.py > /tmp/tables.md
Reads r2_bench.json (N=1, default run), r2_bench_reference.json, r2_bench_n{2,4,8}.json."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return None
    return json.loads(open(p).read().strip().splitlines()[-1])


def kern_row(k):
    return f"| {k['step']} | {k['kernel']} | {k['ms']:.4f} | {k['achieved']:.1f} {k['unit']} | {k.get('frac', float('nan')):.3f} |"


d1 = load("r2_bench.json")
ref = load("r2_bench_reference.json")
byn = {1: d1}
for n in (2, 4, 8):
    byn[n] = load(f"r2_bench_n{n}.json")

print("### Headline: U(1) two-site H_eff apply, D=4096, complex double (BASELINE configs[2])\n")
print("| GPUs | ms / apply (device) | TFLOP/s | efficiency vs N=1 | e2e ms / apply | e2e TFLOP/s | sharded == whole |")
print("|---|---|---|---|---|---|---|")
for n, d in byn.items():
    if d is None:
        continue
    eff = d1["ms_per_step"] / (n * d["ms_per_step"])
    e = d["e2e"]
    print(f"| {n} | {d['ms_per_step']:.3f} | {d['value'] / 1e3:.1f} | {eff:.3f} | {e['ms_per_step']:.3f} | {e['value'] / 1e3:.1f} | "
          f"{d.get('sharded_vs_unsharded_rel_err', '-') if n > 1 else '-'} |")
print()
if ref:
    print(f"Reference arm (`bench.py --impl reference`, same config, {ref['cpu_baseline']['cores']} host threads): "
          f"{ref['ms_per_step']:.0f} ms / apply = {ref['value']:.1f} GFLOP/s; device-timed ratio at N=1 "
          f"{d1['value'] / ref['value']:.0f}x, end-to-end ratio {d1['e2e']['value'] / ref['value']:.0f}x.\n")
r = d1["roofline"]
print(f"Roofline (N=1, {r['kernel']}): {r['achieved']:.2f} TFLOP/s algorithmic / {r['peak']:.2f} measured ZGEMM = {r['frac']:.3f} "
      f"(executed ratio {r['executed_ratio']}, pipe fraction {r['pipe_frac']:.3f}); DRAM traffic per launch {r['traffic'] / 1e9:.2f} GB.\n")
print("| step | kernel | ms | achieved | fraction of roof |")
print("|---|---|---|---|---|")
for k in d1["kernels"]:
    print(kern_row(k))
print()
e = d1["e2e"]
c = e.get("cold_dropin") or {}
print(f"End to end at N=1: pipelined {e['ms_per_step']:.2f} ms, serial {e['serial_ms_per_step']:.2f} ms "
      f"({e['h2d_bytes_per_step'] / 1e6:.0f} MB up, {e['d2h_bytes_per_step'] / 1e6:.0f} MB down per apply); cold drop-in "
      f"(4 x match + plan + execute from pageable host memory) {c.get('ms_per_step', float('nan')):.0f} ms of which match "
      f"{c.get('match_ms', float('nan')):.2f} ms, plan {c.get('plan_ms', float('nan')):.1f} ms.")
f = d1.get("fused_mpo")
if f:
    print(f"Pre-contracted two-site MPO chain: {f['ms_per_step']:.3f} ms / apply (MPO step {f['mpo_step_ms']:.3f} ms, {f['mpo_step_gbs']:.0f} GB/s).")
p = d1.get("parity")
if p:
    print(f"Parity of this run against the reference on the same operands: rel. Frobenius {p['rel_fro_vs_reference']:.2e}, same block structure {p['same_block_structure']}.")
cb = d1["cpu_baseline"]
print(f"cpu_baseline (same config): {cb['value']:.1f} GFLOP/s on {cb['cores']} threads; thread sweep at D={cb['thread_sweep']['D']}: "
      + ", ".join(f"{t}: {v:.1f}" for t, v in sorted(cb['thread_sweep']['gflops_by_threads'].items(), key=lambda x: int(x[0]))) + " GFLOP/s.")
print(f"Clocks during the timed region: {d1['clocks']}; kernel launches in the timed region: {d1['gpu_launches']}.\n")

print("### The other named configurations (sub-records of the same runs)\n")
print("| config | GPUs | ms / apply | TFLOP/s | % of measured FP64 GEMM peak | speed-up vs N=1 | e2e ms | cpu_baseline GFLOP/s (threads) | parity vs reference | sharded == whole |")
print("|---|---|---|---|---|---|---|---|---|---|")
for key in (d1.get("sub_records") or {}):
    base = d1["sub_records"][key]
    for n, d in byn.items():
        if d is None or not d.get("sub_records") or key not in d["sub_records"] or d["sub_records"][key] is None:
            continue
        v = d["sub_records"][key]
        cpu = v.get("cpu_baseline") or {}
        par = (v.get("parity") or {}).get("rel_fro_vs_reference")
        if par is None:
            par = v.get("parity_rel_err_sampled_blocks")
        e2e = (v.get("e2e") or {}).get("ms_per_step")
        cpu_s = f"{cpu['value']:.1f} ({cpu.get('cores')})" if cpu.get("value") else "-"
        par_s = f"{par:.1e}" if par is not None else "-"
        print(f"| {key} | {n} | {v['ms_per_step']:.4g} | {v['value'] / 1e3:.2f} | {v.get('pct_fp64_peak', float('nan')) / (1 if n == 1 else 1):.1f} | "
              f"{base['ms_per_step'] / v['ms_per_step']:.2f} | {e2e if e2e is None else round(e2e, 3)} | {cpu_s} | {par_s} | "
              f"{v.get('sharded_vs_unsharded_rel_err', '-') if n > 1 else '-'} |")
print()
for key, v in (d1.get("sub_records") or {}).items():
    print(f"{key} (N=1), per kernel:\n")
    print("| step | kernel | ms | achieved | fraction of roof |")
    print("|---|---|---|---|---|")
    for k in v.get("kernels", []):
        print(kern_row(k))
    print()
