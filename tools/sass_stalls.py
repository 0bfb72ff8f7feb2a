#!/usr/bin/env python
"""Sum of the static stall counts (control words) of a SASS address range: the issue time of one warp running alone, without
scoreboard waits.  usage: python tools/sass_stalls.py <cuobjdump -sass output> <lo hex> <hi hex>   (see tools/jit_sass.py for the dump)"""
import re,sys,collections
path=sys.argv[1]; lo=int(sys.argv[2],16); hi=int(sys.argv[3],16)
lines=open(path).read().split("\n")
ins=[]
i=0
while i<len(lines):
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/",lines[i])
    if m and i+1<len(lines):
        m2=re.match(r"\s*/\* (0x[0-9a-f]+) \*/",lines[i+1])
        if m2:
            ins.append((int(m.group(1),16),m.group(2).strip(),int(m.group(3),16),int(m2.group(1),16)))
            i+=2; continue
    i+=1
tot=0; n=0; waits=0; h=collections.Counter()
for a,t,w0,w1 in ins:
    if lo<=a<=hi:
        stall=(w1>>41)&0xf; yld=(w1>>45)&1; wmask=(w1>>52)&0x3f
        tot+=stall; n+=1; waits+= 1 if wmask else 0
        h[stall]+=1
print("instr",n,"sum of static stall counts",tot,"instr with scoreboard wait",waits,sorted(h.items()))
