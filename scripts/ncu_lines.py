#!/usr/bin/env python3
"""Executed warp instructions and stall samples of one profiled launch, aggregated by CUDA source line.

ncu's `--page source --csv` lists per-SASS-instruction counters but no line numbers; `nvdisasm -g` of the cubin inside the
shipped library lists the same instructions in the same order WITH their `-lineinfo` source lines.  This script joins the two
(refusing to if the mnemonics do not line up, i.e. the library was rebuilt since the capture) and prints the hottest lines.

  python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [launch_index=0] [top=30]
  e.g. python scripts/ncu_lines.py gpurun_out/r01am/prof_linearize.ncu-rep packet_search 0 40
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(report, kernel, launch):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True, check=True).stdout
    secs, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    # every launch appears once per source view that has data; keep the sections that carry SASS rows
    secs = [s for s in secs if s["hdr"] and "Instructions Executed" in s["hdr"]]
    s = secs[launch]
    h = s["hdr"]
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    return s["name"], [(r[ia].strip(), int(r[ie] or 0), int(r[isamp] or 0)) for r in s["rows"] if len(r) > ie]


def disassembly(kernel):
    lib = os.path.join(ROOT, "small_gicp_b200", "lib", "libsgicp_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
        dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
        starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and re.search(kernel, l)]
        for st in starts:
            cur, seq = None, []
            for l in dis[st + 1 :]:
                if l.startswith("//-----") or l.startswith("\t.section"):
                    break
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    cur = (m.group(1), int(m.group(2)))
                    continue
                m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
                if m:
                    seq.append((cur, m.group(2).strip()))
            yield dis[st], seq


def mnemonic(text):
    t = text.split()
    return (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0] if t else "")[:5]


def main():
    report, kernel = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    name, data = sass_rows(report, kernel, launch)
    match = None
    for label, seq in disassembly(kernel):
        if len(seq) == len(data) and all(mnemonic(a[1]) == mnemonic(b[0]) for a, b in zip(seq, data)):
            match = seq
            break
    if match is None:
        raise SystemExit(f"no function in the library matches the {len(data)} profiled instructions of {name[:60]} (rebuilt since the capture?)")
    agg = collections.defaultdict(lambda: [0, 0])
    for (key, _), (_, ex, sm) in zip(match, data):
        agg[key][0] += ex
        agg[key][1] += sm
    tot, tots = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    print(f"{name[:90]}\nlaunch {launch}: {tot / 1e6:.2f} M warp instructions, {tots} stall samples, {len(data)} SASS instructions")
    cache = {}
    for key, (ex, sm) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        f, ln = key if key else ("?", 0)
        if f not in cache:
            try:
                cache[f] = open(f).read().splitlines()
            except OSError:
                cache[f] = []
        text = cache[f][ln - 1].strip()[:110] if 0 < ln <= len(cache[f]) else ""
        print(f"{100 * ex / tot:5.1f} % instr {100 * sm / max(1, tots):5.1f} % samples  {os.path.basename(f)}:{ln:<4d} {text}")


if __name__ == "__main__":
    main()
