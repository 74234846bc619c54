"""Distance (in SASS instructions) from every LDS.128 of a kernel to the first instruction that reads its result.
usage: python tools/sass_lds_distance.py <library.so> <kernel name substring>
The decoder's chain warp issues in order: a table row that is loaded right before its use stalls the walk for the whole
shared-memory latency (a change in a COLD tier moved these loads and cost 20 %).  Good: >= 10 for the five
loads inside the fast tier of cm_decode_kernel."""
import re,sys,subprocess
so=sys.argv[1]; fn=sys.argv[2]
txt=subprocess.run(["cuobjdump","-sass",so],capture_output=True,text=True).stdout
lines=[];on=False
for l in txt.splitlines():
    if "Function :" in l: on = fn in l
    elif on and re.match(r"\s*/\*[0-9a-f]{4}\*/",l): lines.append(re.sub(r"/\* 0x[0-9a-f]+ \*/","",l).strip())
ins=[re.sub(r"^/\*[0-9a-f]+\*/\s*","",l) for l in lines]
res=[]
for i,s in enumerate(ins):
    m=re.search(r"LDS\.128 R(\d+),",s)
    if not m: continue
    r=int(m.group(1)); regs={f"R{r+k}" for k in range(4)}
    d=None
    for j in range(i+1,min(i+80,len(ins))):
        ops=ins[j].split(None,1)
        body=ins[j]
        # source operands: everything after first comma
        srcs=body.split(",",1)[1] if "," in body else ""
        if any(re.search(r"\b"+x+r"\b(?!\d)",srcs) for x in regs): d=j-i;break
    res.append((i,d))
print(len(ins),"instr; LDS.128 -> first use distance:",[d for _,d in res])
