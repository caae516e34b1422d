"""ncu launch list (--csv --page raw, one row per kernel launch) of `bench.py --no-graph --quick` -> the launches of the LAST
training iteration (profiles/rN_ncu_launches_one_step.csv) and a per-kernel summary (profiles/rN_ncu_launch_list_summary.csv).
An iteration ends with its second adam_ema_kernel launch (D update, then G update).
Usage: python tools/ncu_step_summary.py gpurun_out/launches.csv profiles/r3 ["header note"]"""
import csv
import re
import sys
from collections import OrderedDict

src, prefix = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src, newline="")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
body = [r for r in rows[hi + 2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return 0.0


units = dict(zip(hdr, rows[hi + 1]))


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)


def to_mb(v, unit):
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)


T, R, W, P = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
names = [re.sub(r"^void ", "", r[col["Kernel Name"]]).split("(")[0] for r in body]
adam = [i for i, n in enumerate(names) if n.startswith("adam_ema_kernel")]
assert len(adam) >= 2, "no complete iteration in the capture"
end = adam[-1]
start = adam[-3] + 1 if len(adam) >= 3 else 0
sel = range(start, end + 1)
out = []
for i in sel:
    r = body[i]
    out.append((r[col["ID"]], names[i], r[col["Grid Size"]], r[col["Block Size"]], to_us(num(r, T), units.get(T, "us")),
                to_mb(num(r, R), units.get(R, "byte")), to_mb(num(r, W), units.get(W, "byte")), num(r, P)))
total = sum(o[4] for o in out)
with open(prefix + "_ncu_launches_one_step.csv", "w") as f:
    f.write("# ncu --metrics %s --clock-control none\n" % ",".join((T, R, W, P)))
    f.write("# %s; the launches of ONE training iteration (D update + G update); cold-cache, serialised per-launch times\n" % note)
    f.write("id,kernel,grid,block,time_us,dram_read_MB,dram_write_MB,tensor_pipe_active_pct\n")
    for o in out:
        f.write('%s,"%s","%s","%s",%.2f,%.3f,%.3f,%.1f\n' % o)
agg = OrderedDict()
for o in out:
    a = agg.setdefault(o[1], [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += o[4]; a[2] += o[5]; a[3] += o[6]; a[4] += o[7] * o[4]
with open(prefix + "_ncu_launch_list_summary.csv", "w") as f:
    f.write("# per kernel over ONE training iteration (source: %s_ncu_launches_one_step.csv); %d launches, %.1f ms summed; %s\n"
            % (prefix.split("/")[-1], len(out), total / 1e3, note))
    f.write("kernel,launches,total_us,share_pct,avg_us,dram_read_MB_per_launch,dram_write_MB_per_launch,time_weighted_tensor_pipe_active_pct\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.1f,%.2f,%.1f,%.2f,%.2f,%.1f\n' % (k[:60], a[0], a[1], 100 * a[1] / total, a[1] / a[0], a[2] / a[0], a[3] / a[0],
                                                          a[4] / max(a[1], 1e-9)))
print("%d launches, %.2f ms summed" % (len(out), total / 1e3))
