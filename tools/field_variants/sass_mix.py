#!/usr/bin/env python
"""SASS instruction mix of the main body of the kernels matching a regex (callees after the EXIT are skipped), with
a multiplier-pipe cycle model for B200 (tools/imad_ubench.cu: IMAD.WIDE / IMAD.HI 4 cycles per warp instruction and
SMSP, IMAD / IMAD.X / IMAD.MOV 2) and an ALU-pipe one (2 cycles).   python sass_mix.py <cubin|.o> <regex>"""
import re,collections,subprocess,sys
def count(obj, pat):
    txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
    for fn in re.split(r'\n\s*Function : ', txt)[1:]:
        name=fn.split('\n')[0]
        if not re.search(pat,name): continue
        fn=fn.split(' EXIT ')[0]   # main body only (callees follow the EXIT)
        c=collections.Counter()
        for m in re.finditer(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', fn, re.M):
            op=m.group(1)
            k=op.split('.')[0]
            if k=='IMAD':
                if '.WIDE' in op: k='IMAD.WIDE'
                elif '.HI' in op: k='IMAD.HI'
                elif '.MOV' in op: k='IMAD.MOV'
                elif '.X' in op or '.IADD' in op: k='IMAD.X'
            c[k]+=1
        tot=sum(c.values())
        fma=(c['IMAD.WIDE']+c['IMAD.HI'])*4+(c['IMAD']+c['IMAD.MOV']+c['IMAD.X'])*2
        alu=sum(c[k] for k in ('IADD3','SEL','LOP3','SHF','MOV','PLOP3','ISETP','VIADD'))*2
        print(name[:40],'total',tot,'fma-cycles',fma,'alu-cycles',alu,dict(c.most_common(9)))
count(sys.argv[1], sys.argv[2])
