"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line: samples and the top stall reasons."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rev = sys.argv[3] if len(sys.argv) > 3 else None  # git revision the profiled binary was built from: source text is taken from there
import subprocess
_files = {}
def text_of(fname, line):
    if rev is None: return None
    if fname not in _files:
        try: _files[fname] = subprocess.run(["git", "show", f"{rev}:gubernator_b200/csrc/{fname}"], capture_output=True, text=True, check=True).stdout.splitlines()
        except Exception: _files[fname] = []
    L = _files[fname]
    return L[line - 1].strip() if 0 < line <= len(L) else ""
cur = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None or len(r) < len(hdr): continue
    try: line = int(r[0])
    except ValueError: continue
    key = (cur, line)
    if r[1].strip(): src[key] = r[1].strip()
    i_s = hdr.index("# Samples")
    try: n = int(r[i_s] or 0)
    except ValueError: n = 0
    agg[key]["samples"] += n
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: agg[key][h] += int(r[j] or 0)
            except ValueError: pass
    try: agg[key]["inst"] += int(r[hdr.index("Instructions Executed")] or 0)
    except ValueError: pass
tot = sum(v["samples"] for v in agg.values())
print("total samples", tot)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((k, c) for k, c in v.items() if k.startswith("stall_")), key=lambda kc: -kc[1])[:3]
    print(f"{100*v['samples']/max(tot,1):5.1f}%  {key[0]}:{key[1]:<5} inst={v['inst']:<7} {' '.join(f'{k[6:]}={c}' for k,c in st)}  | {(text_of(*key) if rev and text_of(*key) is not None else src.get(key,''))[:110]}")
