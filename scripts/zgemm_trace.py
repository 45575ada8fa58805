"""Sum a CARC_ZGEMM_TRACE=1 log (stderr lines "zgemm M N K batch opA opB splits ms") by shape:
    CARC_ZGEMM_TRACE=1 python scripts/sweep_bench.py --sizes 8x8 --cpu-max-D 0 2> trace.log; python scripts/zgemm_trace.py trace.log"""
import sys, collections
by = collections.defaultdict(lambda: [0, 0.0])
for line in open(sys.argv[1]):
    f = line.split()
    if len(f) != 9 or f[0] != "zgemm":
        continue
    M, N, K, batch = (int(v) for v in f[1:5])
    key = (M, N, K, batch, int(f[5]), int(f[6]), int(f[7]))
    by[key][0] += 1
    by[key][1] += float(f[8])
total = sum(v[1] for v in by.values())
print("total %.1f ms in %d products, %d shapes" % (total, sum(v[0] for v in by.values()), len(by)))
print("%8s %8s %8s %6s %3s %3s %3s %6s %10s %7s %8s" % ("M", "N", "K", "batch", "opA", "opB", "spl", "calls", "ms", "share", "TFLOP/s"))
for key, (calls, ms) in sorted(by.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    M, N, K, batch = key[:4]
    print("%8d %8d %8d %6d %3d %3d %3d %6d %10.2f %6.1f%% %8.2f" % (*key, calls, ms, 100 * ms / total, 8.0 * M * N * K * batch * calls / (ms * 1e-3) / 1e12 if ms > 0 else 0))
