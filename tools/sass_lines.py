#!/usr/bin/env python3
"""Static SASS instruction count per CUDA source line of one kernel:
    python tools/sass_lines.py <mangled-name substring> [top N]"""
import collections, os, re, subprocess, sys, tempfile
func = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "betse_b200", "libbetse_b200.so")], cwd=d, capture_output=True)
cnt = collections.Counter(); ops = collections.Counter()
for f in os.listdir(d):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", f], cwd=d, capture_output=True, text=True).stdout
    cur = None; line = None
    for ln in out.split("\n"):
        m = re.match(r"\t\.section\t\.text\.(\S+?),", ln)
        if m: cur = m.group(1); continue
        if ln.startswith("\t.section"): cur = None
        if not cur or func not in cur: continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            cnt[line] += 1
            s = m.group(2).split(); ops[(s[1] if s[0].startswith("@") else s[0]).split(".")[0]] += 1
src = {}
print("total", sum(cnt.values()), ops.most_common(12))
for (f, l), n in cnt.most_common(top):
    if f not in src:
        pth = os.path.join(root, "betse_b200", "csrc", f)
        src[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
    print("%-20s %4d %4d  %s" % (f, l, n, src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""))
