"""First-light parity script: CUDA engine vs the C oracle on seeded channel frames."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import fixedpointldpc_b200 as fp
from oracle import pyoracle as po

def frames_for(code, rate, snr_db, count, seed):
    rng = np.random.default_rng(seed)
    snr = 2 * 10 ** (snr_db / 10) * rate
    sigma = np.sqrt(1 / snr)
    llr = 2 * snr * (1 + sigma * rng.standard_normal((count, code.n)))
    return (llr * 16).astype(np.int32)  # truncation toward zero like int(double)

def tables_of(code):
    vdeg, cdeg, vlist, clist = code.tables()
    return po.Tables(code.n, code.m, vdeg, cdeg, vlist, clist)

ok = True
for name, snrs, count in (("wifi", (2.0, 0.5), 48), ("a5", (4.5, 2.0), 32), ("c79", (4.5, 2.0), 32), ("a24", (6.0, 3.0), 6)):
    code = fp.codes.NAMED[name]()
    t = tables_of(code)
    orc = po.Oracle(t)
    rate = fp.codes.INFO_BITS[name] / code.n
    for snr in snrs:
        llr = frames_for(code, rate, snr, count, 1234)
        for precheck in (False, True):
            exp = [orc.decode(x, precheck=precheck) for x in llr]
            for prec in (32, 16, 0):
                dec = fp.Decoder(code, precheck=precheck, precision=prec)
                t0 = time.time()
                out = dec.decode(llr, want_post=True, want_v2c=True)
                dt = time.time() - t0
                bits = fp.unpack_bits(out["bits"], code.n)
                bad = []
                dm = t.cdeg[None, :] > np.arange(t.dc_max)[:, None]
                for f, (it, b, post, edge) in enumerate(exp):
                    if out["iters"][f] != it: bad.append((f, 'iters', int(out["iters"][f]), it)); continue
                    if not (bits[f] == b).all(): bad.append((f, 'bits')); continue
                    if it == 0: continue
                    if not (out["post"][f] == post).all(): bad.append((f, 'post')); continue
                    if not (out["v2c"][f][dm] == edge[dm]).all(): bad.append((f, 'v2c'))
                st = dec.stats()
                print(name, snr, 'precheck', precheck, 'prec', prec, 'frames', count, 'mismatch', len(bad), bad[:3],
                      'iters', np.bincount(out["iters"].clip(0), minlength=31)[[0,1,2,3,30]].tolist(),
                      'fallback', st['fallback_frames'], 'T', st['threads'], st['threads32'], 'slots', st['frames_per_cta'], st['frames_per_cta32'], 'ms %.1f' % (dt*1e3), flush=True)
                ok &= not bad
                dec.close()
print("ALL OK" if ok else "MISMATCHES")
