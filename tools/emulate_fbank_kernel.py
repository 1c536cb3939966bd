"""Lane-level numpy emulation of avex_b200/csrc/fbank.cu phase A (index math only) -- a CPU design check.
Mirrors the CUDA code line by line: v[t][j] registers, IDX16 map, sXf[q*17+t] exchange, pair post-processing."""
import numpy as np, sys
sys.path.insert(0, '.')
from oracle import kaldi_fbank as OF

def IDX16(q): return 4 * (q & 3) + (q >> 2)

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
        v[4*q0], v[4*q0+1], v[4*q0+2], v[4*q0+3] = fft4(v[4*q0], v[4*q0+1], v[4*q0+2], v[4*q0+3])
    return v

rng = np.random.RandomState(0)
# check fft16 alone
x = rng.randn(16) + 1j * rng.randn(16)
V = fft16(x); nat = np.array([V[IDX16(q)] for q in range(16)])
print("fft16 err", np.abs(nat - np.fft.fft(x)).max())

frame = rng.randn(400) * 1000
win = 0.5 * OF.povey_window(400, np.float64)
tw = np.array([[np.exp(-2j*np.pi*t*q/256) for t in range(16)] for q in range(16)])  # [q][t]
tw2 = np.exp(-2j*np.pi*np.arange(136)/512)
v = np.zeros((16, 16), complex); s = 0.0
for t in range(16):
    for j in range(16):
        n = t + 16*j
        if j < 12 or (j == 12 and t < 8):
            v[t][j] = complex(frame[2*n], frame[2*n+1]); s += frame[2*n] + frame[2*n+1]
mu = s / 400.0
for t in range(16):
    for j in range(16):
        n = t + 16*j
        if j < 12 or (j == 12 and t < 8):
            a0, a1 = v[t][j].real - mu, v[t][j].imag - mu
            prev = a0 if n == 0 else frame[2*n-1] - mu
            v[t][j] = complex((a0 - 0.97*prev)*win[2*n], (a1 - 0.97*a0)*win[2*n+1])
sX = np.zeros(272, complex)
for t in range(16):
    r = fft16(v[t])
    for q in range(1, 16): r[IDX16(q)] *= tw[q][t]
    for q in range(16): sX[q*17 + t] = r[IDX16(q)]
u = np.zeros((16, 16), complex)
for t in range(16):
    for tt in range(16): u[t][tt] = sX[t*17 + tt]
sZ = np.zeros(272, complex)
for t in range(16):
    r = fft16(u[t])
    for p in range(16): sZ[t + 16*p] = r[IDX16(p)]
P = np.zeros(272)
for t in range(16):
    for m in range(9):
        k = t + 16*m
        if m < 8 or t == 0:
            A, Bc, W = sZ[k], sZ[(256-k) & 255], tw2[k]
            ex, ey = A.real + Bc.real, A.imag - Bc.imag
            ox, oy = A.imag + Bc.imag, Bc.real - A.real
            tx, ty = W.real*ox - W.imag*oy, W.real*oy + W.imag*ox
            P[k] = (ex+tx)**2 + (ey+ty)**2; P[256-k] = (ex-tx)**2 + (ey-ty)**2
# reference power spectrum
f = frame - frame.mean(); sh = np.concatenate([f[:1], f[:-1]]); f = (f - 0.97*sh) * OF.povey_window(400, np.float64)
ref = np.abs(np.fft.rfft(f, 512))**2
print("power rel err", np.abs(P[:257] - ref).max() / ref.max())
