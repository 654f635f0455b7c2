"""Shared helpers of the parity tests: build the same inputs for the oracle, the emulation harness and
the CUDA engine."""
import numpy as np

from mongeampere_b200 import inputs


from mongeampere_b200.workloads import make_case, load_engine  # noqa: E402,F401  (the workloads live with the product)


def oracle_for(O, case, nthreads=1):
    cfg = case["cfg"]
    orc = O.Oracle(cfg["vx"], cfg["vy"], cfg["tri"], case["abc"], nthreads=nthreads)
    orc.set_points(case["X"])
    return orc


def hessian_rel_err(H_ref, H):
    """max |H - H_ref| relative to the row diagonal."""
    import scipy.sparse as sp
    D = sp.csr_matrix(H_ref - H)
    if D.nnz == 0:
        return 0.0
    diag = np.maximum(np.abs(H_ref.diagonal()), 1e-300)
    rows = np.repeat(np.arange(D.shape[0]), np.diff(D.indptr))
    return float(np.max(np.abs(D.data) / diag[rows]))


def same_pattern(H_ref, H):
    A = H_ref.copy(); A.sort_indices()
    B = H.copy(); B.sort_indices()
    return A.nnz == B.nnz and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)


def _exact_clip(poly, a, b, c):
    """keep a x + b y + c > 0 (exact)"""
    out=[]; n=len(poly)
    for q in range(n):
        p0,p1=poly[q],poly[(q+1)%n]
        s0=a*p0[0]+b*p0[1]+c; s1=a*p1[0]+b*p1[1]+c
        if s0>0:
            out.append(p0)
            if not s1>0:
                t=s0/(s0-s1); out.append((p0[0]+t*(p1[0]-p0[0]), p0[1]+t*(p1[1]-p0[1])))
        elif s1>0:
            t=s0/(s0-s1); out.append((p0[0]+t*(p1[0]-p0[0]), p0[1]+t*(p1[1]-p0[1])))
    return out

def exact_cell_mass(cfg, X, w, i, cand):
    """Mass of the Laguerre cell of Dirac i on a grid config (inputs.config kind "grid" on [-1,1]^2) in EXACT rational
    arithmetic: the box clipped by the bisectors with every candidate site, the polygon clipped by every overlapped
    triangle, the linear density integrated exactly.  The arbiter when engine and oracle disagree at the 1e-10 level."""
    from fractions import Fraction as Fr
    import bisect
    F=lambda v: Fr(float(v))
    n=cfg["n"]; m=cfg["m"]
    xi,yi,wi=F(X[i,0]),F(X[i,1]),F(w[i])
    poly=[(Fr(-1),Fr(-1)),(Fr(1),Fr(-1)),(Fr(1),Fr(1)),(Fr(-1),Fr(1))]
    for j in cand:
        if j==i: continue
        xj,yj,wj=F(X[j,0]),F(X[j,1]),F(w[j])
        # pow_i < pow_j  <=>  2 x.(yj - yi) ... : a x + b y + c > 0 with
        a=2*(xi-xj); b=2*(yi-yj); c=-xi*xi-yi*yi+xj*xj+yj*yj+wi-wj
        poly=_exact_clip(poly,a,b,c)
        if not poly: return Fr(0)
    # grid squares overlapped
    vx=cfg["vx"]; vy=cfg["vy"]; rho=cfg["rho"]
    gx=[F(vx[a*m]) for a in range(n)]; gy=[F(vy[b]) for b in range(m)]
    xs=[p[0] for p in poly]; ys=[p[1] for p in poly]
    a0=max(bisect.bisect_right(gx,min(xs))-1,0); a1=min(bisect.bisect_left(gx,max(xs)),n-1)
    b0=max(bisect.bisect_right(gy,min(ys))-1,0); b1=min(bisect.bisect_left(gy,max(ys)),m-1)
    mass=Fr(0)
    for a in range(a0,a1):
        for b in range(b0,b1):
            P00=(gx[a],gy[b]); P10=(gx[a+1],gy[b]); P11=(gx[a+1],gy[b+1]); P01=(gx[a],gy[b+1])
            r00,r10,r11,r01=F(rho[a*m+b]),F(rho[(a+1)*m+b]),F(rho[(a+1)*m+b+1]),F(rho[a*m+b+1])
            for tri,rv in (((P00,P10,P11),(r00,r10,r11)),((P00,P11,P01),(r00,r11,r01))):
                pc=poly
                for e in range(3):
                    p,q=tri[e],tri[(e+1)%3]
                    # left of p->q: (q-p) x (x-p) > 0
                    aa=-(q[1]-p[1]); bb=(q[0]-p[0]); cc=-(aa*p[0]+bb*p[1])
                    pc=_exact_clip(pc,aa,bb,cc)
                    if not pc: break
                if not pc or len(pc)<3: continue
                # plane through tri
                (x0,y0),(x1,y1),(x2,y2)=tri; f0,f1,f2=rv
                det=(x1-x0)*(y2-y0)-(x2-x0)*(y1-y0)
                A=((f1-f0)*(y2-y0)-(f2-f0)*(y1-y0))/det; B=((x1-x0)*(f2-f0)-(x2-x0)*(f1-f0))/det; C=f0-A*x0-B*y0
                for k in range(1,len(pc)-1):
                    (ax,ay),(bx,by),(cx,cy)=pc[0],pc[k],pc[k+1]
                    ar=((bx-ax)*(cy-ay)-(cx-ax)*(by-ay))/2
                    mass+=ar*(A*(ax+bx+cx)/3+B*(ay+by+cy)/3+C)
    return mass

