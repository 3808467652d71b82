#!/usr/bin/env python
"""Encoded stall counts of the longest loop of the kernels matching a regex (control bits 41-44 of the upper 64-bit word of
every SASS instruction): instructions, sum and average of the stall cycles, histogram, average on the wide multiplies.
  python tools/field_variants/sass_stalls.py build/csrc/msm_g2.o k_msm_accumulate"""
import re,subprocess,sys,collections
def analyze(obj, pat):
    txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
    for fn in re.split(r'\n\s*Function : ', txt)[1:]:
        name=fn.split('\n')[0]
        if not re.search(pat,name): continue
        lines=fn.split('\n')
        ins=[]
        i=0
        while i<len(lines):
            m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/',lines[i])
            if m and i+1<len(lines):
                m2=re.match(r'\s+/\* (0x[0-9a-f]+) \*/',lines[i+1])
                if m2:
                    ins.append((int(m.group(1),16),m.group(2),int(m2.group(1),16)))
                    i+=2; continue
            i+=1
        best=None
        for a,t,w in ins:
            mm=re.search(r'BRA\s+(0x[0-9a-f]+)',t)
            if mm and int(mm.group(1),16)<a and (best is None or a-int(mm.group(1),16)>best[1]-best[0]): best=(int(mm.group(1),16),a)
        loop=[(a,t,w) for a,t,w in ins if best[0]<=a<=best[1]]
        st=[(w>>41)&0xf for a,t,w in loop]
        n=len(loop)
        hist=collections.Counter(st)
        wide=[((w>>41)&0xf) for a,t,w in loop if 'IMAD.WIDE' in t or 'IMAD.HI' in t]
        print(name[:70]); print('  loop instr',n,'sum stall',sum(st),'avg %.3f'%(sum(st)/n),'hist',dict(sorted(hist.items())),'avg stall on wide %.2f'%(sum(wide)/max(1,len(wide))))
analyze(sys.argv[1], sys.argv[2])
