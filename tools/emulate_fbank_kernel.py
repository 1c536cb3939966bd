"""Lane-level numpy emulation of avex_b200/csrc/fbank.cu (v2) -- a CPU design check of the index plan, in float64.

Mirrors the CUDA kernel step by step for one 16-frame chunk: interleaved staging sD[i] = (d[i], d[i+160]) with partial sums,
per-frame mean from 100 partial sums, 16 x 16 Cooley-Tukey with v[t][j] registers / IDX16 map / xg[q*17+t] transposition, partner
exchange by lane (16 - t) & 15 with the lane-0 special case, ELL mel rows owned by lane t (bins t + 16 i), and the patch-operand
column mapping.  Compared with oracle/kaldi_fbank.py.  Not part of the product or the tests' oracle."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import kaldi_fbank as OF  # noqa: E402

HOP, WIN, FPC = 160, 400, 16
SEG = (FPC - 1) * HOP + WIN
NPAIR = SEG - HOP
NQ = NPAIR // 4


def IDX16(q):
    return 4 * (q & 3) + (q >> 2)


def fft4(a0, a1, a2, a3):
    t0, t1, t2, t3 = a0 + a2, a0 - a2, a1 + a3, a1 - a3
    return t0 + t2, t1 - 1j * t3, t0 - t2, t1 + 1j * t3


def fft16(v):
    v = list(v)
    for j0 in range(4):
        v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12] = fft4(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12])
    C1, S1, R = 0.92387953251128675613, 0.38268343236508977173, 0.70710678118654752440
    v[5] *= complex(C1, -S1); v[9] *= complex(R, -R); v[13] *= complex(S1, -C1)
    v[6] *= complex(R, -R); v[10] = complex(v[10].imag, -v[10].real); v[14] *= complex(-R, -R)
    v[7] *= complex(S1, -C1); v[11] *= complex(-R, -R); v[15] *= complex(-C1, S1)
    for q0 in range(4):
        v[4 * q0], v[4 * q0 + 1], v[4 * q0 + 2], v[4 * q0 + 3] = fft4(v[4 * q0], v[4 * q0 + 1], v[4 * q0 + 2], v[4 * q0 + 3])
    return v


def mel_tables(mel):  # avexk_fbank_create
    start, length = np.zeros(128, int), np.zeros(128, int)
    for j in range(128):
        nz = np.nonzero(mel[:, j])[0]
        if len(nz):
            start[j], length[j] = nz[0], nz[-1] - nz[0] + 1
    mel_len, mel_off, rows, off = [], [], [], 0
    for i in range(8):
        L = max(1, length[16 * i : 16 * i + 16].max())
        mel_len.append(L); mel_off.append(off)
        blk = np.zeros((L, 16))
        for t in range(16):
            j = 16 * i + t
            blk[: length[j], t] = mel[start[j] : start[j] + length[j], j]
        rows.append(blk); off += L
    return start, mel_len, mel_off, np.concatenate(rows)


def chunk(x, T, s0, win, mel, F, prescale=32768.0):
    """x: one clip (float64), chunk starting at sample s0 -> (log-mel [16, 128], patch rows [8, 256])."""
    ld = lambda s: x[s] if 0 <= s < T else 0.0  # noqa: E731
    sD = np.zeros((NPAIR, 2)); sPS = np.zeros(SEG // 4)
    pe = 0.97 * prescale
    for u in range(NQ):
        sa, sb = s0 + 4 * u, s0 + 4 * u + HOP
        A = [ld(sa + k) for k in range(4)]; B = [ld(sb + k) for k in range(4)]
        pa = ld(sa - 1) if sa > 0 else A[0]
        pb = ld(sb - 1)
        pA, pB = [pa] + A[:3], [pb] + B[:3]
        for k in range(4):
            sD[4 * u + k] = (A[k] * prescale - pe * pA[k], B[k] * prescale - pe * pB[k])
        sPS[u] = sum(A) * prescale
        if u >= NQ - HOP // 4:
            sPS[u + HOP // 4] = sum(B) * prescale
    tw1 = np.array([[np.exp(-2j * np.pi * t * q / 256) for t in range(16)] for q in range(16)])  # [q][t]
    tw2 = np.exp(-2j * np.pi * np.arange(136) / 512)
    start, mel_len, mel_off, melw = mel_tables(mel)
    out = np.zeros((16, 128)); patch = np.zeros((8, 256))
    for g in range(8):
        flA = 2 * g
        mu = np.array([sPS[40 * flA : 40 * flA + 100].sum(), sPS[40 * flA + 40 : 40 * flA + 140].sum()]) / 400.0
        dc = -0.03 * mu
        P = np.zeros((272, 2))
        for fr in range(2):  # the two halves of every packed value
            v = np.zeros((16, 16), complex)
            for t in range(16):
                for j in range(16):
                    n = t + 16 * j
                    if j < 12 or (j == 12 and t < 8):
                        d = sD[HOP * flA + 2 * n : HOP * flA + 2 * n + 2, fr]
                        v[t][j] = complex((d[0] + dc[fr]) * win[2 * n], (d[1] + dc[fr]) * win[2 * n + 1])
            xg = np.zeros(16 * 17, complex)
            for t in range(16):
                V = fft16(v[t])
                for q in range(16):
                    xg[q * 17 + t] = V[IDX16(q)] * tw1[q][t]
            Z = np.zeros((16, 16), complex)  # Z[t][p] = register IDX16(p) of lane t = Z[t + 16 p]
            for t in range(16):
                V = fft16([xg[t * 17 + tt] for tt in range(16)])
                for p in range(16):
                    Z[t][p] = V[IDX16(p)]
            for t in range(16):
                src = (16 - t) & 15
                r = [Z[src][8 + i] for i in range(8)]
                for p in range(8):
                    A = Z[t][p]
                    if p == 0:
                        Bc = A if t == 0 else r[7]
                    else:
                        Bc = r[8 - p] if t == 0 else r[7 - p]
                    k = t + 16 * p
                    W = tw2[k]
                    ex, ey, ox, oy = A.real + Bc.real, A.imag - Bc.imag, A.imag + Bc.imag, Bc.real - A.real
                    tx, ty = W.real * ox - W.imag * oy, W.real * oy + W.imag * ox
                    P[k, fr] = (ex + tx) ** 2 + (ey + ty) ** 2
                    P[256 - k, fr] = (ex - tx) ** 2 + (ey - ty) ** 2
                if t == 0:
                    A = Z[0][8]
                    P[128, fr] = (2 * A.real) ** 2 + (2 * A.imag) ** 2
        for t in range(16):
            for i in range(8):
                j = t + 16 * i
                acc = np.zeros(2)
                for m in range(mel_len[i]):
                    acc += melw[mel_off[i] + m, t] * P[start[j] + m]
                val = np.log(np.maximum(acc, np.finfo(np.float32).eps))
                out[flA, j], out[flA + 1, j] = val
                odd = t & 1
                e = (out[flA + 1, j - 1], val[1]) if odd else (val[0], None)
                # patch operand: row i, col = frame * 16 + bin-within-patch
                patch[i, flA * 16 + t] = val[0]
                patch[i, (flA + 1) * 16 + t] = val[1]
    return out, patch


if __name__ == "__main__":
    rng = np.random.RandomState(0)
    T = 16 * 160 * 2 + 400 + 77
    x = rng.randn(T) * 0.1
    F = OF.frame_count(T)
    ref = OF.fbank(x[None] * 32768.0, n_mels=128, dtype=np.float64)[0]
    win = 0.5 * OF.povey_window(400, np.float64)
    mel = OF.mel_filterbank(dtype=np.float64) if "dtype" in OF.mel_filterbank.__code__.co_varnames else OF.mel_filterbank().astype(np.float64)
    worst = 0.0
    for c in range((F + 15) // 16):
        out, patch = chunk(x, T, c * 16 * HOP, win, mel, F)
        n = min(16, F - 16 * c)
        err = np.abs(out[:n] - ref[16 * c : 16 * c + n]).max()
        worst = max(worst, err)
        if n == 16:  # patch mapping: row fp, col i*16+j <-> fbank[tp*16+i, fp*16+j]
            want = ref[16 * c : 16 * c + 16].reshape(16, 8, 16).transpose(1, 0, 2).reshape(8, 256)
            worst = max(worst, np.abs(patch - want).max())
        print(f"chunk {c}: frames {n}, max|err| vs float64 oracle {err:.3e}")
    print("worst", worst)
    assert worst < 1e-6
