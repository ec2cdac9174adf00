#!/usr/bin/env python3
"""profiles/<round>/<tag>_sass.md: mnemonic census of libsgicp_b200.so / libsgicp_b200_prof.so (cuobjdump -sass) + the TMA excerpt.
usage: python scripts/sass_census.py profiles/r02/c_sass.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dst = sys.argv[1]
out = ["# %s -- SASS census of the two libraries (`cuobjdump -sass`, sm_100a only)\n" % os.path.basename(dst).split("_")[0],
       "`libsgicp_b200.so` is the product (production kernels, no environment switches); `libsgicp_b200_prof.so` is the same sources built with `-DSGB_PROFILING`",
       "(superseded A/B kernels + `SGB_*` switches), loaded only by the A/B scripts and `test_search_structures_agree`.\n",
       "| mnemonic | what it is | product | profiling |", "|---|---|---|---|"]
what = {"UBLKCP": "`cp.async.bulk` (TMA 1-D bulk copy): leaf staging A/B of the packet search -- measured slower, profiling library only",
        "SYNCS": "mbarrier arrive / try_wait of that bulk copy", "UTMALDG": "tensor-map TMA (not used: nothing here is a tile)",
        "LDGSTS": "`cp.async` (factor kernel's two-tiles-ahead operand pipeline)", "ACQBULK": "`griddepcontrol.wait` (programmatic dependent launch of kernels 2-4)",
        "REDUX": "warp-wide integer min / max (`__reduce_min_sync`: nearest point of a leaf / a ring in one instruction)",
        "DFMA": "FP64 fused multiply-add (factor algebra, sums)", "SHFL": "warp shuffles (transposing reduction, query broadcast)", "VOTE": "warp votes (divergence-free packet walk)",
        "LDG.E.ENL2.256": "256-bit global loads (`ld.global.nc.v8.f32`, new with sm_100): one pair record of a block list per load (probe, ring search)",
        "FFMA2": "packed FP32 fma (`fma.rn.f32x2`, new with sm_100): both squared distances of a pair record at once", "FADD2": "packed FP32 add (query - point, two points)",
        "FMUL2": "packed FP32 multiply",
        "ATOMG": "global atomics (work queues, pending / class lists, tickets)", "MEMBAR": "fences (ticket tree, peer mailboxes)", "LDG.E.128": "128-bit global loads (float4 SoA streams)"}
sass = {}
for lib in ("libsgicp_b200.so", "libsgicp_b200_prof.so"):
    sass[lib] = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "small_gicp_b200", "lib", lib)], capture_output=True, text=True).stdout
for m, w in what.items():
    out.append(f"| `{m}` | {w} | {sass['libsgicp_b200.so'].count(m)} | {sass['libsgicp_b200_prof.so'].count(m)} |")
archs = sorted(set(re.findall(r"sm_\d+a?", subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "small_gicp_b200", "lib", "libsgicp_b200.so")], capture_output=True, text=True).stdout)))
out.append(f"\nFunctions: {sass['libsgicp_b200.so'].count('Function :')} (product), {sass['libsgicp_b200_prof.so'].count('Function :')} (profiling); ELF images: " + ", ".join(archs) + ".\n")
s = sass["libsgicp_b200_prof.so"]
i = s.find("packet_search_kernelILb1E")
seg = s[i : s.find("Function :", i + 10)]
lines = [l for l in seg.splitlines() if re.search(r"UBLKCP|SYNCS", l)][:10]
out.append("## `packet_search_kernel<true, 10>` (TMA leaf staging, profiling library): the bulk copy and its mbarrier\n\n```")
out += [re.sub(r"\s+/\*\s*0x[0-9a-f]+\s*\*/", "", l).rstrip() for l in lines]
out.append("```\n")
names = sorted(set(re.findall(r"Function : (\S+)", sass["libsgicp_b200.so"])))
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
short = sorted(set(re.sub(r"\(.*", "", d).replace("void ", "") for d in dem if "sgb::" in d))
out.append("## Kernels of the product library (own code; the rest is `cub::DeviceRadixSort` / `DeviceScan`)\n\n" + ", ".join(f"`{x}`" for x in short) + "\n")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
