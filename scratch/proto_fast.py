"""numpy prototype of the restructured (flavour-basis, traceless Cayley-Hamilton) propagation."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle

def fast_probs(dm, mix, mat_pot, lri, nubar, E, rho, dist, variant="trig"):
    N, L = rho.shape
    D = np.diag([0, dm[1,0], dm[2,0]]).astype(complex)
    Hv = mix @ D @ mix.conj().T          # flavour basis, nu
    sv = 1.0 if nubar > 0 else -1.0      # nubar: H' = -Hvac/2E + aV + lri (overall conj irrelevant)
    inv2E = 0.5 / E
    Tprod = None
    for l in range(L):
        d = dist[:, l]; r = rho[:, l]
        a = 0.5 * r * 1.52588e-4
        h = sv * Hv[None] * inv2E[:, None, None] + a[:, None, None] * mat_pot[None] + (lri * 1e9)[None]
        h = 0.5 * (h + h.conj().transpose(0, 2, 1))
        h2 = h @ h
        d0, d1, d2 = (np.real(h[:,i,i]) for i in range(3))
        n01 = np.abs(h[:,0,1])**2; n02 = np.abs(h[:,0,2])**2; n12 = np.abs(h[:,1,2])**2
        c2 = -(d0+d1+d2)
        c1 = d0*(d1+d2) + d1*d2 - n01 - n12 - n02
        rpa = np.real(h[:,0,1]*h[:,1,2]*h[:,2,0])
        c0 = d0*n12 + d1*n02 + d2*n01 - 2*rpa - d0*d1*d2
        p = np.maximum(c2*c2 - 3*c1, 0)
        q = -13.5*c0 - c2**3 + 4.5*c1*c2
        tmp = np.maximum(27*(0.25*c1*c1*(p-c1) + c0*(q+6.75*c0)), 0)
        b_ = (2/3)*np.sqrt(p)
        if variant == "cbrt":
            zr, zi = q, np.sqrt(tmp)
            n2 = zr*zr + zi*zi
            inv = 1/np.sqrt(np.where(n2 > 0, n2, 1.0))
            zr, zi = np.where(n2 > 0, zr*inv, 1.0), np.where(n2 > 0, zi*inv, 0.0)
            th0 = (np.arctan2(zi.astype(np.float32), zr.astype(np.float32)) / np.float32(3)).astype(np.float32)
            # emulate fast intrinsics: add ~5e-7 abs noise
            cf = np.cos(th0).astype(np.float32) + np.float32(4e-7); sf = np.sin(th0).astype(np.float32) - np.float32(3e-7)
            c, s_ = cf.astype(np.float64), sf.astype(np.float64)
            m = c*c + s_*s_ - 1.0
            r = 1 - 0.5*m + 0.375*m*m
            c, s_ = c*r, s_*r
            for _ in range(1):
                w2r, w2i = c*c - s_*s_, 2*c*s_
                w3r, w3i = w2r*c - w2i*s_, w2r*s_ + w2i*c
                e_ = zi*w3r - zr*w3i
                d_ = e_/3
                k = 1 - 0.5*d_*d_
                c, s_ = c*k - s_*d_, s_*k + c*d_
            cth, sth = c, s_
        else:
            th = np.arctan2(np.sqrt(tmp), q)/3
            cth, sth = np.cos(th), np.sin(th)
        c120, s120 = -0.5, np.sqrt(3)/2
        lam = np.stack([b_*cth, b_*(cth*c120 - sth*s120), b_*(cth*c120 + sth*s120)], axis=1) - (c2/3)[:,None]
        I2 = None
        t = 2*2.534*d
        e = np.exp(-1j*lam*t[:,None])
        l0,l1,l2 = lam[:,0],lam[:,1],lam[:,2]
        Dk = np.stack([(l0-l1)*(l0-l2),(l1-l0)*(l1-l2),(l2-l0)*(l2-l1)],axis=1)
        sm = np.stack([l1+l2,l0+l2,l0+l1],axis=1); pr = np.stack([l1*l2,l0*l2,l0*l1],axis=1)
        w = e / Dk
        a2 = w.sum(axis=1)
        a1 = -(w*sm).sum(axis=1)
        a0 = (w*pr).sum(axis=1)
        T = a0[:,None,None]*np.eye(3)[None] + a1[:,None,None]*h + a2[:,None,None]*h2
        active = d > 0
        T = np.where(active[:,None,None], T, np.eye(3)[None].astype(complex))
        Tprod = T if Tprod is None else T @ Tprod
    P = np.abs(Tprod)**2      # P[n, j, i] = |T[j,i]|^2 = prob i->j
    return P.transpose(0, 2, 1)

if __name__ == "__main__":
    g = np.load("tests/golden/ref_prob3_f8.npz")
    prem = np.loadtxt("pisa_b200/resources/osc/PREM_12layer.dat")
    Lr = oracle.OracleLayers(prem, 2.0, 20.0); Lr.setElecFrac(0.4656, 0.4656, 0.4957)
    _, den, dis = Lr.calcLayers(g["coszen"])
    keys = sorted({k.rsplit("/", 1)[0] for k in g.files if k.count("/") == 2})
    for key in keys:
        ref = g[key + "/probability"]
        out = fast_probs(g[key+"/dm"], g[key+"/mix"], g[key+"/mat_pot"], g[key+"/lri_pot"], int(g[key+"/nubar"]), g["energy"], den, dis, variant=sys.argv[1] if len(sys.argv)>1 else "trig")
        err = np.abs(out - ref)
        rel = err / np.maximum(np.abs(ref), 1e-300)
        ok = np.isclose(out, ref, rtol=1e-10, atol=1e-14)
        print("%-40s max abs %.2e  max rel %.2e  frac outside(1e-10,1e-14) %.2e  unit %.2e" % (key, err.max(), rel.max(), 1-ok.mean(), np.abs(out.sum(axis=2)-1).max()))
