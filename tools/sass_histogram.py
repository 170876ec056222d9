#!/usr/bin/env python
"""Static SASS opcode histogram of the main kernels in libb2sense.so (cuobjdump -sass): how many
instructions of each opcode the compiled kernel CONTAINS (not executed counts -- those are in the ncu
summaries).  Shows at a glance that the cast is fp64 / integer code with no tensor-core (HMMA / UTCMMA)
or TMA (UBLKCP) instructions, and how much of it is local-memory traffic (LDL / STL).
usage: sass_histogram.py [lib] > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "mujoco_ros2_simulation_b200/libb2sense.so"
KERNELS = ["collect_kernelILi0ELb1E", "drain_kernelILi0ELb1E", "mesh_pairs_kernelILi0E", "mesh_pairs_coop_kernelILi0E", "mesh_ids_kernel",
           "mesh_final_kernelILi0E", "cast_kernelILi0ELb1ELb1ELb0E", "refit_kernel", "depth_to_cloud_kernel",
           "rgb_flip_kernel"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            hist.setdefault(cur, collections.Counter())[m.group(1)] += 1
    print(f"# static SASS opcode counts, {LIB} (sm_100a, nvcc -fmad=false -lineinfo)")
    for key in KERNELS:
        for name, h in hist.items():
            if key in name:
                tot = sum(h.values())
                fp64 = sum(v for k, v in h.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"))
                loc = h.get("LDL", 0) + h.get("STL", 0)
                tc = sum(v for k, v in h.items() if k.startswith(("HMMA", "IMMA", "DMMA", "UTC", "UBLKCP", "UTMA")))
                print(f"\n{name}\n  {tot} instructions; fp64 arithmetic {fp64} ({100*fp64/tot:.1f} %), local ld/st {loc} "
                      f"({100*loc/tot:.1f} %), tensor-core / TMA {tc}")
                print("  " + "  ".join(f"{k}:{v}" for k, v in h.most_common(28)))
                break


if __name__ == "__main__":
    main()
