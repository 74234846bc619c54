"""Per-kernel SASS comparison of two builds of the library (cuobjdump -sass, instruction text without addresses):
which kernels are machine-code identical, which changed, which are new.  Used to show that the kernels that run by
default are byte-for-byte the ones that passed the GPU suite, whatever was added around them:

    git archive <commit> | tar -x -C /tmp/old && (cd /tmp/old && nvcc <flags of bzip3_b200/build.py> -o old.so bzip3_b200/csrc/bz3_api.cu)
    python tools/sass_compare.py /tmp/old/old.so bzip3_b200/libbzip3_b200.so
"""
import subprocess, re, hashlib, sys
def funcs(so):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    res = {}
    name = None
    body = []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name: res[name] = body
            name = m.group(1); body = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            # keep the instruction text only (strip address comment and encoding)
            ins = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()
            body.append(ins)
    if name: res[name] = body
    return res
a = funcs(sys.argv[1]); b = funcs(sys.argv[2])
print(len(a), len(b))
same = diff = 0
for k in sorted(a):
    if k in b:
        ha = hashlib.md5("\n".join(a[k]).encode()).hexdigest(); hb = hashlib.md5("\n".join(b[k]).encode()).hexdigest()
        if ha == hb: same += 1
        else:
            diff += 1; print("DIFF", k[:110], len(a[k]), len(b[k]))
    else:
        print("GONE", k[:110])
print("same", same, "diff", diff, "new", len([k for k in b if k not in a]))
