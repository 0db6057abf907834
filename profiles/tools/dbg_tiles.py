import numpy as np, sys
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
a = a[(a[:, 0] > 0) & (a[:, 5] > 0)]
t0 = a[:, 0].min()
print("tiles", len(a), "kernel span us", (a[:, 5].max() - t0) / 1e3)
d = lambda i, j: (a[:, j] - a[:, i]) / 1e3
for name, i, j in [("start->scan done", 0, 1), ("scan->lookback done", 1, 2), ("scan->worker0 staged", 1, 3),
                   ("scan->barrier released", 1, 4), ("barrier->flush done", 4, 5), ("total", 0, 5)]:
    x = d(i, j)
    print("%-26s mean %.2f  p50 %.2f  p90 %.2f  p99 %.2f us" % (name, x.mean(), np.median(x), np.percentile(x, 90), np.percentile(x, 99)))
# concurrency: tiles in flight over time
starts = np.sort(a[:, 0]); ends = np.sort(a[:, 5])
mid = (t0 + a[:, 5].max()) // 2
print("in flight at mid:", (a[:, 0] <= mid).sum() - (a[:, 5] <= mid).sum())
