import sys, numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.float64)
a = a[a[:,1] > 0]
t0 = a[:,1].min()
b = a[:,1] - t0; e = a[:,2] - t0
print("ctas", len(a), "begin ns min/med/max", b.min(), np.median(b), b.max(), "end ns min/med/max", e.min(), np.median(e), e.max())
print("groups per cta min/med/max", a[:,4].min(), np.median(a[:,4]), a[:,4].max(), "sum", a[:,4].sum())
print("particles per cta min/med/max", a[:,5].min(), np.median(a[:,5]), a[:,5].max(), "sum", a[:,5].sum())
dur = e - b
rate = a[:,5] / np.maximum(dur, 1)
print("ns per particle per cta: min/med/max", (1/rate[rate>0]).min(), np.median(1/rate[rate>0]), (1/rate[rate>0]).max())
# per SM
sm = a[:,3].astype(int)
cnt = np.bincount(sm)
print("ctas per SM: min/max", cnt[cnt>0].min(), cnt.max(), "SMs used", (cnt>0).sum())
for q in (0, 10, 50, 90, 100): print("end pct", q, np.percentile(e, q))
order = np.argsort(e)
print("earliest finishing:", a[order[:5]][:, [0,3,4,5]].tolist(), e[order[:5]].tolist())
print("latest finishing:", a[order[-5:]][:, [0,3,4,5]].tolist(), e[order[-5:]].tolist())
