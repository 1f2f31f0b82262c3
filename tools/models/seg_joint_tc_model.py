"""Python model of csrc/seg_joint_tc.cu::seg_joint_tc_kernel (forward joint on tcgen05): the SAME work decomposition, index
formulas (Toeplitz operand address = (x + v)*64 + c*4, accumulator block / partial / reduce layout, touched mask) and
ring-slot mbarrier protocol (producer / consumer run as coroutines; a stall of both is a deadlock), checked against the
oracle's F.conv2d formulation.  It validates everything about the kernel except the UMMA descriptor semantics, which
tools/umma_sw64_probe.cu checks on hardware.   python tools/models/seg_joint_tc_model.py"""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import seg_losses as oseg
SJ_U, SJ_MT, SJ_SLOTS, SJ_ST2 = 7, 3, 8, 2

def plan(n, h, w, T, sms=148):
  V = 2*T+1; wp = (w+15)//16*16; ug = (V+SJ_U-1)//SJ_U
  want = max(1, (3*sms + n*ug - 1)//(n*ug)); yc = (h+want-1)//want; yc = max(yc, T+1); yc = min(yc, h)
  return dict(V=V, wp=wp, ugroups=ug, ychunk=yc, nychunks=(h+yc-1)//yc)

class Bar:
  def __init__(s): s.phase = 0  # number of completed phases
  def done(s, parity): return (s.phase & 1) != parity  # try_wait.parity(p) succeeds iff phase with parity p completed
def sim_joint(x1m, x2m, T):
  n, h, w, KP = x1m.shape; P = plan(n, h, w, T); V, wp = P["V"], P["wp"]
  items = n*P["nychunks"]; ctas = items*P["ugroups"]
  part = np.full((ctas, SJ_U, 24, 16, 16), np.nan)
  for bid in range(ctas):
    g = bid % P["ugroups"]; yc = (bid//P["ugroups"]) % P["nychunks"]; img = bid//(P["ugroups"]*P["nychunks"])
    ya, yb = yc*P["ychunk"], min(h, yc*P["ychunk"]+P["ychunk"]); u0 = g*SJ_U; nu = min(SJ_U, V-u0)
    r_lo = max(0, ya+u0-T); r_hi = min(h-1, (yb-1)+(u0+nu-1)-T)
    # ---- protocol simulation: producer ops list and consumer ops list executed round-robin until both finish
    full1=[Bar() for _ in range(SJ_SLOTS)]; empty1=[Bar() for _ in range(SJ_SLOTS)]; full2=[Bar() for _ in range(SJ_ST2)]; empty2=[Bar() for _ in range(SJ_ST2)]
    slot_row=[None]*SJ_SLOTS; st2_row=[None]*SJ_ST2
    def producer():
      r_next = r_lo
      for y in range(ya, yb):
        need = min(r_hi, y+(u0+nu-1)-T)
        while r_next <= need:
          idx = r_next-r_lo; s = idx % SJ_SLOTS
          while not empty1[s].done(((idx//SJ_SLOTS)&1)^1): yield
          slot_row[s] = r_next; full1[s].phase += 1; r_next += 1
        it = y-ya; s2 = it % SJ_ST2
        while not empty2[s2].done(((it//SJ_ST2)&1)^1): yield
        st2_row[s2] = y; full2[s2].phase += 1
      return
    acc = np.zeros((SJ_U, SJ_MT, 128, 16)); touched = 0
    def consumer():
      nonlocal touched
      for y in range(ya, yb):
        it = y-ya; s2 = it % SJ_ST2
        while not full2[s2].done((it//SJ_ST2)&1): yield
        assert st2_row[s2] == y
        b = np.zeros((wp, 16)); b[:w] = x2m[img, y]
        for ul in range(nu):
          r = y+u0+ul-T
          if r < 0 or r >= h: continue
          idx = r-r_lo; s = idx % SJ_SLOTS
          while not full1[s].done((idx//SJ_SLOTS)&1): yield
          assert slot_row[s] == r, (slot_row[s], r)
          buf = np.zeros((wp+24, 16)); xs = np.arange(wp+24)-T; ok = (xs>=0)&(xs<w); buf[ok] = x1m[img, r, xs[ok]]
          flat = buf.reshape(-1)
          for mt in range(SJ_MT):
            for kk in range(wp//16):
              for xk in range(16):
                x = kk*16+xk
                # A[m][k=xk] = flat[(kk*16+mt*8)*16 + xk*16 + m]  (m = vl*16+c; M atom stride == K-row stride == one pixel)
                a = flat[(kk*16+mt*8+xk)*16:(kk*16+mt*8+xk)*16+128]
                acc[ul, mt] += np.outer(a, b[x])
            touched |= 1 << (ul*SJ_MT+mt)
          if ul == 0: empty1[s].phase += 1
        empty2[s2].phase += 1
      return
    pr, co = producer(), consumer(); pd = cd = False; stall = 0
    while not (pd and cd):
      progressed = False
      if not pd:
        try: next(pr)
        except StopIteration: pd = True; progressed = True
      if not cd:
        try: next(co)
        except StopIteration: cd = True; progressed = True
      stall = 0 if progressed else stall+1
      assert stall < 10000, ("deadlock", bid)
    # epilogue
    t2 = 0
    for ul in range(nu):
      lo, hi = max(ya, T-u0-ul), min(yb, h+T-u0-ul)
      if lo < hi: t2 |= 7 << (ul*SJ_MT)
    assert t2 == touched, (bin(t2), bin(touched))
    for blk in range(SJ_U*SJ_MT):
      ul, mt = blk//SJ_MT, blk % SJ_MT
      live = (t2 >> blk) & 1
      for m in range(128):
        vl, c = m >> 4, m & 15
        part[bid, ul, mt*8+vl, c, :] = acc[ul, mt, m] if live else 0.0
  k = KP
  joint = np.zeros((V*V, 16, 16))
  for u in range(V):
    for v in range(V):
      g, ul = u//SJ_U, u % SJ_U
      for it in range(items): joint[u*V+v] += part[it*P["ugroups"]+g, ul, v]
  return joint

def main():
  rng = np.random.default_rng(0)
  for (n, h, T) in [(2, 12, 3), (1, 20, 10), (2, 16, 5), (1, 9, 4)]:
    w = h
    x1m = rng.random((n, h, w, 16)); x2m = rng.random((n, h, w, 16))
    got = sim_joint(x1m, x2m, T)
    A = oseg.seg_joint_displacements(torch.from_numpy(x1m).permute(0,3,1,2), torch.from_numpy(x2m).permute(0,3,1,2), T)
    want = A.permute(2,3,0,1).reshape(-1,16,16).numpy()
    print((n,h,T), plan(n,h,w,T), "max err", np.abs(got-want).max(), "scale", np.abs(want).max())


if __name__ == "__main__":
  main()
