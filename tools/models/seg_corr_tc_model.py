"""Python model of csrc/seg_joint_tc.cu::seg_corr_tc_kernel (backward contractions on tcgen05): work decomposition,
coefficient-block layout Bg[u][j][c][c'], input-row ring protocol and output partials, checked against a direct
evaluation of  out[Y,X,c] = scale * sum H[u][v][c][c'] in[Y - s(u-T), X - s(v-T), c'].
   python tools/models/seg_corr_tc_model.py"""
import sys, numpy as np
SC_U, SC_SLOTS = 6, 7
class Bar:
  def __init__(s): s.phase = 0
  def done(s, parity): return (s.phase & 1) != parity
def plan(n, h, T, sms=148):
  V=2*T+1; ug=(V+SC_U-1)//SC_U; want=max(1,(3*sms+n*ug-1)//(n*ug)); yc=(h+want-1)//want; yc=max(yc,4); yc=min(yc,h)
  return V, ug, yc, (h+yc-1)//yc
def ref(inp, H, T, sgn, scale):
  n,h,w,KP = inp.shape; V=2*T+1; out=np.zeros_like(inp)
  for u in range(V):
    for v in range(V):
      for Y in range(h):
        yr = Y - sgn*(u-T)
        if yr<0 or yr>=h: continue
        for X in range(w):
          xr = X - sgn*(v-T)
          if xr<0 or xr>=w: continue
          out[:,Y,X,:] += scale * inp[:,yr,xr,:] @ H[u*V+v].T   # out[c] = sum_c' H[c][c'] in[c']
  return out
def sim(inp, H, T, sgn, scale):
  n,h,w,KP = inp.shape; V, ug, ychunk, nyc = plan(n,h,T)
  # Bg[u][j][c][c'] : v = 2T-j (sgn>0) else j
  Bg = np.zeros((V,V,16,16))
  for u in range(V):
    for j in range(V):
      v = 2*T-j if sgn>0 else j
      Bg[u,j] = H[u*V+v]*scale
  part = np.zeros((ug,n,h,w,16))
  for bid in range(n*nyc*ug):
    g = bid % ug; yc=(bid//ug)%nyc; img=bid//(ug*nyc)
    Ya,Yb = yc*ychunk, min(h, yc*ychunk+ychunk); u0=g*SC_U; nu=min(SC_U, V-u0)
    d_first=-sgn*(u0-T); d_last=-sgn*(u0+nu-1-T); dmin,dmax=min(d_first,d_last),max(d_first,d_last)
    r_lo=max(0,Ya+dmin); r_hi=min(h-1,Yb-1+dmax)
    full=[Bar() for _ in range(SC_SLOTS)]; empty=[Bar() for _ in range(SC_SLOTS)]; slot_row=[None]*SC_SLOTS
    def producer():
      r_next=r_lo
      for Y in range(Ya,Yb):
        need=min(r_hi, Y+dmax)
        while r_next<=need:
          idx=r_next-r_lo; s=idx%SC_SLOTS
          while not empty[s].done(((idx//SC_SLOTS)&1)^1): yield
          slot_row[s]=r_next; full[s].phase+=1; r_next+=1
    def consumer():
      for Y in range(Ya,Yb):
        acc=np.zeros((128,16)); anyv=False
        for ul in range(nu):
          d=-sgn*(u0+ul-T); r=Y+d
          if r<0 or r>=h: continue
          idx=r-r_lo; s=idx%SC_SLOTS
          while not full[s].done((idx//SC_SLOTS)&1): yield
          assert slot_row[s]==r
          buf=np.zeros((128+V-1,16)); xs=np.arange(128+V-1)-T; ok=(xs>=0)&(xs<w); buf[ok]=inp[img,r,xs[ok]]
          for j in range(V):
            # A[X][(c')] = buf[X + j][c'] ; B[c][c'] = Bg[u0+ul][j][c][c']
            acc += buf[j:j+128] @ Bg[u0+ul, j].T
          if d==dmin: empty[s].phase+=1
          anyv=True
        part[g,img,Y,:,:] = acc[:w] if anyv else 0.0
    pr,co=producer(),consumer(); pd=cd=False; stall=0
    while not (pd and cd):
      prog=False
      if not pd:
        try: next(pr)
        except StopIteration: pd=True; prog=True
      if not cd:
        try: next(co)
        except StopIteration: cd=True; prog=True
      stall=0 if prog else stall+1
      assert stall<10000, ("deadlock", bid)
  return part.sum(0)
def main():
  rng=np.random.default_rng(1)
  for (n,h,T) in [(2,12,3),(1,24,10),(1,9,3)]:
    for sgn in (1,-1):
      V=2*T+1
      inp=rng.random((n,h,h,16)); H=rng.standard_normal((V*V,16,16)); H=(H+H.transpose(0,2,1))/2
      a=sim(inp,H,T,sgn,0.37); b=ref(inp,H,T,sgn,0.37)
      print((n,h,T,sgn), plan(n,h,T), "max err", np.abs(a-b).max(), "scale", np.abs(b).max())


if __name__ == "__main__":
  main()
