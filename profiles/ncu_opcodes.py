"""Dynamic SASS opcode mix of a kernel from an `ncu --set full --import-source on` report:
   python profiles/ncu_opcodes.py REPORT.ncu-rep UNITS [--json OUT]
UNITS = work units of the captured launch (check-node updates), so that the table reads "warp instructions per unit"."""
import collections, csv, io, json, subprocess, sys

rep, units = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hd = None
ops = collections.Counter()
for r in rows:
    if len(r) > 2 and r[0] == "Address":
        hd = r
        continue
    if hd and len(r) == len(hd):
        try:
            n = int(r[hd.index("Instructions Executed")])
        except ValueError:
            continue
        t = r[hd.index("Source")].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        ops[op.split(".")[0]] += n
tot = sum(ops.values())
table = {k: round(v / units, 1) for k, v in ops.most_common(40)}
print("warp instructions per unit: %.1f" % (tot / units))
for k, v in table.items():
    print("%8.1f  %5.1f %%  %s" % (v, 100 * v * units / tot, k))
if "--json" in sys.argv:
    json.dump({"report": rep.split("/")[-1], "units": units, "warp_instructions_per_unit": tot / units, "opcodes_per_unit": table},
              open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
