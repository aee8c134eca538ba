import numpy as np,sys
t=np.load(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/tc_trace.npy').astype(np.float64)
n=int((t[:,9]!=0).sum()); t=t[:n]; lo,hi=n//4,3*n//4; seg=t[lo:hi]
print(n,'tiles; cycles/tile',(seg[-1,9]-seg[0,9])/(len(seg)-1))
def d(a,b,shift=0):
    x=seg[:,b]-seg[:,a] if shift==0 else seg[shift:,b]-seg[:-shift,a]
    return f'{np.mean(x):7.0f} (p10 {np.percentile(x,10):6.0f} p90 {np.percentile(x,90):6.0f})'
print('P: top->wait gfree done     ',d(2,3))
print('P: h part                   ',d(3,4))
print('P: wait xfree               ',d(0,1))
print('P: x copies+land -> next top',d(1,2,1))
print('I: tfree wait (prev issue->)',d(7,5,1))
print('I: hfull wait               ',d(5,6))
print('I: h MMAs + xfull wait      ',d(6,10))
print('I: x MMAs issue             ',d(10,7))
print('E: wait done                ',d(8,9))
print('E: gates(+stores)           ',d(9,12))
print('E: head                     ',d(12,13))
print('E: team cycle (2 tiles)     ',d(8,8,2))
print('xfree -> next same-stage done', d(12,9,2))
t0=t[lo,2]
for i in range(lo,lo+6): print(i,' '.join(f'{(v-t0):7.0f}' for v in t[i,:14]))
