"""Condenses gpurun_out/sanitizer_<tool>.log into profiles/r2_sanitizer.summary.txt (error summary lines, the distinct error
kinds with their first occurrence, and the pytest result line of each tool)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = []
for tool in ("memcheck", "racecheck", "initcheck_tma1", "initcheck_tma0"):
    path = os.path.join(ROOT, "gpurun_out", f"sanitizer_{tool}.log")
    if not os.path.isfile(path):
        continue
    txt = open(path, errors="replace").read().splitlines()
    out.append(f"== compute-sanitizer --tool {tool.split('_')[0]}{' DOST_GEMM_TMA_EPI=' + tool[-1] if '_tma' in tool else ''}  (scripts/gpu_sanitize.sh)")
    kinds = {}
    for i, ln in enumerate(txt):
        m = re.match(r"=+ (Invalid|Uninitialized|Race|Error|Program hit|Potential|Warning)[^\n]*", ln)
        if m:
            key = re.sub(r"0x[0-9a-f]+|\d+", "N", ln)[:160]
            kinds.setdefault(key, (ln, txt[i + 1:i + 4]))
    for key, (ln, ctx) in list(kinds.items())[:12]:
        out.append("  " + ln.strip()[:200])
        for c in ctx:
            out.append("      " + c.strip()[:200])
    for ln in txt:
        if "ERROR SUMMARY" in ln or "RACECHECK SUMMARY" in ln or "Device Frame" in ln and "Uninit" in "".join(txt[max(0, txt.index(ln) - 6):txt.index(ln)]) or re.search(r"\d+ (passed|failed)", ln) or ln.startswith("exit "):
            out.append("  " + ln.strip()[:200])
open(os.path.join(ROOT, "profiles", "r2_sanitizer.summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
