"""Offline statistics (numpy, CPU) of EXACT necessary conditions for `response > 15` on a bench frame: what share of
the pixels / of the 8-pixel row cells each candidate pre-filter of the cascade kernel (chess_cascade.cu) would let
through. Used to decide which filters are worth writing as CUDA; see DESIGN.md, section K1.
    python tools/filter_stats.py"""
import numpy as np, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from mrgingham_b200 import synth
W,H=3840,2160
img=synth.board_frame(W,H,10,seed=0).astype(np.int32)
off=[(2,-5),(0,-5),(-2,-5),(-4,-4),(-5,-2),(-5,0),(-5,2),(-4,4),(-2,5),(0,5),(2,5),(4,4),(5,2),(5,0),(5,-2),(4,-4)]
y0,y1,x0,x1=7,H-7,7,W-7
s=[img[y0+dy:y1+dy, x0+dx:x1+dx] for dx,dy in off]
N=(y1-y0)*(x1-x0)
def frac(m): return 100.0*m.sum()/N
def cells(m):   # 8-pixel row cells aligned to x multiples of 8 (absolute x)
    pad=np.zeros((m.shape[0], W), bool); pad[:, x0:x1]=m
    return 100.0*pad.reshape(m.shape[0], W//8, 8).any(2).sum()/(m.shape[0]*W//8)
A=lambda i,j: np.abs(s[i]-s[j])
# current L1: pixels X..X+3 chords: (0,4),(9,5),(10,6),(7,11); pixels X+4..X+7: (8,12),(9,13),(2,14),(7,11)
xs=np.arange(x0,x1)
first=((xs%8)<4)[None,:]
cur_a=A(0,4)+A(9,5)+A(10,6)+A(7,11)
cur_b=A(8,12)+A(9,13)+A(2,14)+A(7,11)
cur=np.where(first,cur_a,cur_b)>=8
print("current L1: pixels %.2f%% cells %.2f%%"%(frac(cur),cells(cur)))
# alternative: quadrant-redundant chords (i, i+4) for i=0..3
alt=(A(0,4)+A(1,5)+A(2,6)+A(3,7))>=8
print("chords (i,i+4): pixels %.2f%% cells %.2f%%"%(frac(alt),cells(alt)))
alt2=(A(0,4)+A(1,5)+A(2,6)+A(3,7))>=8
# two independent chord sets ANDed
altB=(A(4,8)+A(5,9)+A(6,10)+A(7,11))>=8
print("AND of (i,i+4) and (i+4,i+8): pixels %.2f%% cells %.2f%%"%(frac(alt&altB),cells(alt&altB)))
altC=(A(8,12)+A(9,13)+A(10,14)+A(11,15))>=8
altD=(A(12,0)+A(13,1)+A(14,2)+A(15,3))>=8
m4=alt&altB&altC&altD
print("AND of all four chord sets: pixels %.2f%% cells %.2f%%"%(frac(m4),cells(m4)))
# min of two adjacent chords per i
mn=sum(np.minimum(A(i,i+4),A(i+4,i+8)) for i in range(4))>=8
print("sum_i min(|a-b|,|b-c|): pixels %.2f%% cells %.2f%%"%(frac(mn),cells(mn)))
mn4=sum(np.minimum(np.minimum(A(i,i+4),A(i+4,i+8)),np.minimum(A(i+8,i+12),A(i+12,i))) for i in range(4))>=8
print("sum_i min of 4 chords: pixels %.2f%% cells %.2f%%"%(frac(mn4),cells(mn4)))
# L2 bound
T=sum(A(i,i+4)+A(i+8,i+12)-A(i,i+8)-A(i+4,i+12) for i in range(4))
l2=T>=16
print("L2 (T>=16): pixels %.3f%% cells %.3f%%"%(frac(l2),cells(l2)))
# exact separation bound: sum_i 2*max(min(a,c)-max(b,d), min(b,d)-max(a,c))
sep=sum(2*np.maximum(np.minimum(s[i],s[i+8])-np.maximum(s[i+4],s[i+12]), np.minimum(s[i+4],s[i+12])-np.maximum(s[i],s[i+8])) for i in range(4))
print("sum term_i >=16: pixels %.3f%% cells %.3f%%"%(frac(sep>=16),cells(sep>=16)))
# half-L2: only diameters as a veto: sum chords(cur) - sum of 2 diameters?
# cheap veto idea: sum_i |a-c| (4 diameters) large => edge. response>15 needs sum_i (|a-b|+|c-d|) >= 16 + sum_i(|a-c|+|b-d|)
U=sum(A(i,i+4)+A(i+8,i+12) for i in range(4)); V=sum(A(i,i+8)+A(i+4,i+12) for i in range(4))
for k in (1,2):
    pass
# L1.5 candidates evaluated on pixels flagged by current L1:
for name,m in (("min2",mn),("min4",mn4),("and4",m4)):
    both=cur&m
    print("current L1 & %s: pixels %.2f%% cells %.2f%%"%(name,frac(both),cells(both)))
