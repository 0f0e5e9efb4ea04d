"""Pretty-print the op / GEMM-shape breakdown of a bench.py --profile-ops JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2))
ops = d["op_ms"]
tot = 0.0
for k, v in sorted(((k, v) for k, v in ops.items() if k != "_by_shape"), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]
    print(f"{k:18s} n={v['n_per_step']:5.0f}  {v['ms_per_step']:7.3f} ms")
print("sum", round(tot, 2))
print("-- by shape")
for k, v in sorted(ops.get("_by_shape", {}).items(), key=lambda kv: -kv[1]["us_each"] * kv[1]["n_per_step"])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v['us_each']*v['n_per_step']:8.1f} us  n={v['n_per_step']:.0f}  {k}")
print("-- gemm shapes (M,N,K,a_mode,e_mode)")
AM = ["RAW", "AFF", "AFFS", "SILU", "GN", "BNB"]; EM = ["STORE", "SILU", "SILU_BWD", "GN_BWD"]
gs = sorted(d["gemm_shapes"], key=lambda r: -r["us_each"] * r["n_per_step"])
t = 0
for r in gs[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    M, N, K, a, e = r["M,N,K,a_mode,e_mode"]
    t += r["us_each"] * r["n_per_step"]
    print(f"{r['us_each']*r['n_per_step']:8.1f} us  n={r['n_per_step']:.0f} each {r['us_each']:7.1f}  M={M:8d} N={N:4d} K={K:4d} {AM[a]:5s} {EM[e]:8s} {r['GBps']:7.0f} GB/s")
print("gemm total us", round(sum(r["us_each"] * r["n_per_step"] for r in gs), 1))
