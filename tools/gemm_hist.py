"""Aggregate the per-launch GEMM CSV (bench.py --profile-csv) by size class."""
import sys, csv, collections
rows = list(csv.DictReader(open(sys.argv[1])))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    M, N, K = int(r["M"]), int(r["N"]), int(r["K"])
    key = (int(r["tc"]), min(M, N), K)
    a = agg[key]; a[0] += 1; a[1] += float(r["ms"]); a[2] += float(r["useful_flop"])
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.1f} ms over {len(rows)} launches")
print("tc  min(M,N)      K   launches       ms   share   TF/s(useful)")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{key[0]:2d} {key[1]:9d} {key[2]:6d} {a[0]:10d} {a[1]:8.2f} {100*a[1]/tot:6.1f}% {a[2]/a[1]/1e9 if a[1] else 0:9.1f}")
