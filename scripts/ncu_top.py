"""Dev: top stall lines of an ncu report's source page.  usage: ncu_top.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print('total samples', tot, 'warp instr', sum(f(r, 'Instructions Executed') for r in data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print('stall mix:', {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:N]:
    st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
    print(f"{f(r,'# Samples'):8.0f} {100*f(r,'# Samples')/tot:5.1f}% ex={f(r,'Instructions Executed'):9.0f}  {r[ix['Source']][:64]:64s} {[(int(a), b[6:]) for a, b in st]}")
