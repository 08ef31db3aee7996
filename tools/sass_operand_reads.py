#!/usr/bin/env python3
"""Counts, per basic block of a kernel's SASS, the register source operands that have to be fetched from the register
file (RZ, immediates, constant-bank and uniform operands, predicates and `.reuse` operands excluded; an aligned
64-bit pair counts once with W64=1, twice with W64=2).  Development tool behind the operand-fetch model in
profiles/r01_poseidon_v6_experiments.md.   usage: W64=1 python tools/sass_operand_reads.py <binary|.so> <mangled kernel>"""
import sys,re,subprocess,collections,os
W64=int(os.environ.get("W64","2"))
def body(binf,sym):
    txt=subprocess.run(['cuobjdump','-sass','-fun',sym,binf],capture_output=True,text=True).stdout
    ins=[]
    for l in txt.splitlines():
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)\s*(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(3),m.group(4)))
    return ins
def reads(op,args):
    a=[x.strip() for x in args.split(',')] if args.strip() else []
    regs=[]
    seen_dest=False
    for x in a:
        x0=x.lstrip('-~|!')
        if re.match(r'^(U?P\d|U?PT)$',x0): continue
        m=re.match(r'^R(\d+)(\.\w+)*$',x0)
        if m or x0=='RZ' or x0.startswith('RZ.'):
            if not seen_dest: seen_dest=True; continue
            if x0.startswith('RZ'): continue
            reuse='.reuse' in x0
            regs.append((int(m.group(1)),reuse))
        elif not seen_dest and re.match(r'^UR\d+',x0): seen_dest=True
        elif 'c[' in x0 or 'desc[' in x0:
            mm=re.search(r'\[R(\d+)',x0)
            if mm: regs.append((int(mm.group(1)),False))
    b=op.split('.')[0]
    n=0
    for i,(r,reuse) in enumerate(regs):
        w=1
        if b in("DFMA","DADD","DMUL"): w=W64
        if b=="IMAD" and "WIDE" in op and i==2: w=W64
        if reuse: continue
        n+=w
    return n
if __name__=='__main__':
    binf,sym=sys.argv[1],sys.argv[2]
    ins=body(binf,sym)
    # split into blocks by branches; report per block
    blocks=[[]]
    for adr,op,args in ins:
        if op.split('.')[0] in('BRA','EXIT'): blocks.append([])
        else: blocks[-1].append((op,args))
    for i,b in enumerate(blocks):
        if len(b)<100: continue
        tot=sum(reads(o,a) for o,a in b)
        nre=sum(a.count('.reuse') for o,a in b)
        print(f"block {i}: {len(b)} instrs, {tot} register-word reads ({tot/len(b):.2f}/instr), {nre} reuse operands")
