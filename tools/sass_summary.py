#!/usr/bin/env python3
"""Per-kernel instruction evidence from the shipped library (no GPU needed):
    python tools/sass_summary.py > profiles/sass_summary.txt
Counts the SASS mnemonics that prove the Blackwell-native paths (tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA ->
UTMALDG/UTMASTG, 1-D bulk copies -> UBLKCP, cp.async -> LDGSTS, legacy tensor path -> HMMA) and the resource usage
(`cuobjdump -sass -res-usage`)."""
import os
import re
import subprocess
import sys

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(REPO, "deep-fem-uav-wing_b200", "lib", "libdfw_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "ATOM", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = " ".join(line.split())
            cur = None
    kernels = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {k: 0 for k in MNEMONICS}
            kernels[cur]["_n"] = 0
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_n"] += 1
            for k in MNEMONICS:
                if op == k or (k in ("LDG", "STG", "LDS", "STS", "ATOM", "RED", "SYNCS") and op.startswith(k)) or (k.startswith("UTC") and op.startswith(k)):
                    kernels[cur][k] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    head = subprocess.run(["git", "-C", REPO, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"libdfw_b200.so, sm_100a SASS (cuobjdump -sass / -res-usage), source tree at commit {head}")
    print("columns: instructions | tcgen05.mma (UTC*MMA) | tcgen05.commit (UTCBAR) | tcgen05.ld (LDTM) | tcgen05.st (STTM) | TMA load (UTMALDG) | TMA store (UTMASTG) | "
          "bulk copy (UBLKCP) | cp.async (LDGSTS) | mbarrier ops (SYNCS*) | legacy HMMA | resources\n")
    for (mangled, c), name in sorted(zip(kernels.items(), demangle), key=lambda t: t[1]):
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("dfw::", "")
        mma = c["UTCHMMA"] + c["UTCQMMA"]
        print(f"{short[:58]:58s} {c['_n']:6d} | {mma:3d} | {c['UTCBAR']:3d} | {c['LDTM']:3d} | {c['STTM']:3d} | {c['UTMALDG']:3d} | {c['UTMASTG']:3d} | {c['UBLKCP']:3d} | "
              f"{c['LDGSTS']:3d} | {c['SYNCS']:3d} | {c['HMMA']:3d} | {usage.get(mangled, '')}")


if __name__ == "__main__":
    main()
