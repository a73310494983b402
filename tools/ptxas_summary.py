"""Registers / spills per kernel instantiation from a build/*.ptxas.log:  python tools/ptxas_summary.py fused_uni [filter]"""
import re, subprocess, sys, os
name = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
log = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multigridcmt_b200", "build", name + ".cu.ptxas.log")).read()
for b in re.split(r"ptxas info\s+: Compiling entry function", log)[1:]:
    mangled = re.search(r"'(\S+)'", b).group(1)
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(mgcmt::LevelDev.*", "", dem).replace("void mgcmt::", "")
    if flt and flt not in dem:
        continue
    regs = re.search(r"Used (\d+) registers", b).group(1)
    sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b).groups()
    print("%-70s regs=%s spill=%s/%s" % (dem, regs, sp[0], sp[1]))
