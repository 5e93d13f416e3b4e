"""registers / spills of the render kernels from the ptxas logs next to the objects"""
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else "render"
for f in ("kernels.ptxas.log", "kernels_exact.ptxas.log"):
    text = subprocess.run(["c++filt"], input=open(f"/root/repo/bendy_tracer_b200/csrc/{f}").read(), capture_output=True, text=True).stdout
    name, sp = None, "?"
    for l in text.splitlines():
        m = re.search(r"Compiling entry function '(.*)' for", l)
        if m:
            k = re.search(r"(\w+<[^>]*>)\(", m.group(1))
            name = k.group(1) if k else m.group(1)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", l)
        if m:
            sp = m.group(2)
        m = re.search(r"Used (\d+) registers", l)
        if m and name and pat in name:
            print(f"{f[:14]:15s} {name:50s} {m.group(1):>4s} regs  spill {sp}")
