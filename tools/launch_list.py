"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> per-kernel table:  python tools/launch_list.py in.csv out.md "title" """
import csv
import sys
from collections import OrderedDict

src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 6]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = OrderedDict()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    t = float(r[iv].replace(",", ""))
    t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    a = acc.setdefault(r[ik], [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in acc.values())
with open(out, "w") as f:
    f.write("# %s\n\n" % title)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` -- per-launch times are cold-cache and "
            "serialised; compare SHARES.\nTemplate arguments of push_stream_kernel: <real, threads, unroll, MODE (1 kick, "
            "2 final + next stage-0 deposit, 3 init, 4 kick with the stage-0 drift redone on load), deposit, exact_w, interp, gather through the texture pipe>.\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k[:120], n, t, 100 * t / tot))
print(open(out).read())
