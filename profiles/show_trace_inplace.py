import numpy as np,sys
t=np.load(sys.argv[1]).astype(np.float64)
sh=int(sys.argv[2]) if len(sys.argv)>2 else 2   # epilogue's next tile: 2 (two teams) or 1 (one team)
n=int((t[:,9]!=0).sum()); t=t[:n]; lo,hi=n//4,3*n//4; seg=t[lo:hi]
print(n,'tiles; cycles/tile',(seg[-1,9]-seg[0,9])/(len(seg)-1))
def d(a,b,shift=0):
    x=seg[:,b]-seg[:,a] if shift==0 else seg[shift:,b]-seg[:-shift,a]
    return f'{np.mean(x):7.0f} (p10 {np.percentile(x,10):6.0f} p50 {np.percentile(x,50):6.0f} p90 {np.percentile(x,90):6.0f})'
print('P: top -> hfree wait done    ',d(2,3))
print('P: h stores + arrive         ',d(3,4))
print('P: own loads issue           ',d(4,0))
print('P: wait xfree                ',d(0,1))
print('P: x issue + split -> next top',d(1,2,1))
print('I: prev commit -> gfree       ',d(7,5,1))
print('I: gfree -> xfull             ',d(5,6))
print('I: x MMAs issue + hfull wait  ',d(6,10))
print('I: h MMAs issue + commit      ',d(10,7))
print('E: top -> done                ',d(8,9))
print('E: gates                      ',d(9,11))
print('E: stores                     ',d(11,12))
print('E: head                       ',d(12,13))
print('E: cycle                      ',d(8,8,sh))
print('gates end -> next same-stage done', d(11,9,2))
print('stores end(hfree) -> hfull arrive', d(12,4,2))
print('issuer h commit(7) -> epi done(9)', d(7,9))
t0=t[lo,2]
for i in range(lo,lo+6): print(i,' '.join(f'{(v-t0):7.0f}' for v in t[i,:14]))
