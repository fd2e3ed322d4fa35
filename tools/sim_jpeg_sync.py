"""How fast does a baseline-JPEG Huffman parse started at an arbitrary bit fall in step with the true parse?  (design input of K-J2s,
retto_b200/csrc/jpeg_decode.cu): parse one bench page (no restart markers) from every 8192-bit boundary as if an MCU started there and
report the distance to the first MCU start it shares with the true parse.   python tools/sim_jpeg_sync.py"""
import io, sys, numpy as np, re
from PIL import Image
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.synth import gen_page
u=gen_page(4,1280,1280)[0]
b=io.BytesIO(); Image.fromarray(u).save(b,"JPEG",quality=90,subsampling=2)
d=b.getvalue()
# parse DHT
pos=2; tabs={}
while pos < len(d):
    assert d[pos]==0xFF; m=d[pos+1]; L=(d[pos+2]<<8)|d[pos+3]
    if m==0xC4:
        q=pos+4; end=pos+2+L
        while q<end:
            tc=d[q]>>4; th=d[q]&15; bits=list(d[q+1:q+17]); n=sum(bits); vals=list(d[q+17:q+17+n]); q+=17+n
            code=0; k=0; lut={}
            for l in range(1,17):
                for i in range(bits[l-1]):
                    lut[(l,code)]=vals[k]; code+=1; k+=1
                code<<=1
            tabs[(tc,th)]=lut
    if m==0xDA: sos=pos+2+L; break
    pos+=2+L
ecs=d[sos:]
# unstuff, drop EOI
out=bytearray(); i=0
while i<len(ecs):
    c=ecs[i]
    if c==0xFF:
        if ecs[i+1]==0: out.append(0xFF); i+=2; continue
        else: break
    out.append(c); i+=1
bits=np.unpackbits(np.frombuffer(bytes(out),np.uint8))
N=len(bits); print("clean bytes", len(out), "bits", N)
def sym(tab, p):
    code=0
    for l in range(1,17):
        if p+l>N: return None
        code=(code<<1)|int(bits[p+l-1])
        if (l,code) in tab: return l, tab[(l,code)]
    return 16, 0
def parse(p, b, kk, pend, nb=6, nb0=4):
    """returns list of MCU start positions and final state"""
    starts=[]
    while True:
        if b==0 and kk==0:
            if p>=pend: break
            starts.append(p)
        elif p>=pend: break
        ci = 0 if b<nb0 else 1
        if kk==0:
            r=sym(tabs[(0,ci)],p)
            if r is None: break
            l,s=r; p+=l+s; kk=1
        else:
            r=sym(tabs[(1,ci)],p)
            if r is None: break
            l,s=r; rr=s>>4; sz=s&15; p+=l+sz
            if sz: kk+=rr+1
            else: kk = kk+16 if rr==15 else 64
        if kk>=64:
            kk=0; b+=1
            if b==nb: b=0
    return starts,(p,b,kk)
true_starts,_=parse(0,0,0,N)
print("true MCUs", len(true_starts))
ts=set(true_starts)
SUB=8192
fails=0; dists=[]
for k in range(1, N//SUB):
    r0=k*SUB
    st,_=parse(r0,0,0,min(r0+4*SUB,N))
    common=[x for x in st if x in ts]
    if not common: fails+=1; dists.append(None); continue
    dists.append(common[0]-r0)
print("subs", len(dists), "never sync within 4 regions", fails)
dd=[x for x in dists if x is not None]
print("sync distance bits: median", np.median(dd), "p90", np.percentile(dd,90), "max", max(dd), " >8192:", sum(1 for x in dd if x>8192))
