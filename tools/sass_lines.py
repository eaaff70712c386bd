"""Static SASS instruction count per source line of one entry point (nvdisasm -g -c output on stdin or a cubin path).
Usage: python tools/sass_lines.py <cubin> <function> [file-substring]"""
import subprocess, sys, re, collections
cubin, fn = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
sec = txt.split("//--------------------- .text.%s " % fn)[1].split("//--------------------- ")[0]
cur = None
cnt = collections.Counter()
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        cnt[cur] += 1
print("total", sum(cnt.values()))
for (f, l), c in sorted(cnt.items()):
    if flt in f:
        print(f"{f}:{l}\t{c}")
