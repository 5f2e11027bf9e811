"""Print ptxas register / spill numbers per kernel: python scripts/regs.py [filter]"""
import re, subprocess, sys
out = subprocess.run([sys.executable, "-m", "pygmu2_b200.build", "--force", "-v"], capture_output=True, text=True)
t = out.stdout + out.stderr
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores[^\n]*\n[^\n]*Used (\d+) registers", t):
    n = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    if flt in n:
        print(f"{n[:70]:70s} regs {m.group(4):>3s} stack {m.group(2)} spill {m.group(3)}")
if out.returncode: print(t[-3000:])
