"""CPU study (no GPU): local error of CF4 / Richardson-CF4 / CF6 on the C2 sweep as a function of step length.
Dense N<=10 model of H(t) = Dint - delta(t) N_r + (Omega(t)/2) X with the C2 register statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.linalg import expm
from pulser_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spec = W.config_c2(n=n, seed=20)
U = spec.interaction_matrix[0]
D = 1 << n
idx = np.arange(D)
bits = np.array([(idx >> (n - 1 - k)) & 1 for k in range(n)])  # digit of qubit k ; eigenbasis r,g => digit 0 = r
nr = (bits == 0).astype(float)
dint = np.zeros(D)
for i in range(n):
    for j in range(i + 1, n):
        dint += U[i, j] * nr[i] * nr[j]
Nr = nr.sum(0)
X = np.zeros((D, D))
for k in range(n):
    X[idx, idx ^ (1 << (n - 1 - k))] += 1.0
omega = 2 * np.pi * 1.5
U0 = omega / 2
d0, df = -6 * U0, 2 * U0
def delta(t):  # t in us, sweep part only (t measured from sweep start)
    return d0 + (df - d0) * t / 2.5
def Hm(t):
    return np.diag(dint - delta(t) * Nr) + 0.5 * omega * X
def expmH(A):  # exp(-i A), A hermitian
    w, v = np.linalg.eigh(A)
    return (v * np.exp(-1j * w)) @ v.conj().T
# exact moments for linear-in-time H on [a,b]: B0 = h H(mid), B1 = (h^2/12) H'
Hp = np.diag(-(df - d0) / 2.5 * Nr)
def cf4(a, b):
    h = b - a; mid = 0.5 * (a + b)
    B0 = h * Hm(mid); B1 = (h * h / 12.0) * Hp
    return expmH(0.5 * B0 + 2 * B1) @ expmH(0.5 * B0 - 2 * B1)
def fine(a, b, m=64):
    ts = np.linspace(a, b, m + 1)
    Uu = np.eye(D, dtype=complex)
    for i in range(m):
        Uu = cf4(ts[i], ts[i + 1]) @ Uu
    return Uu
# a state on the trajectory: evolve from all-ground through part of the sweep with fine steps
psi = np.zeros(D, complex); psi[D - 1] = 1.0
# rise (500 ns) approx: skip, start the sweep with Omega on from ground (non adiabatic but populated similarly)
t0 = 1.2
Uu = fine(0.0, t0, 600)
psi = Uu @ psi
print("n", n, "norm", np.linalg.norm(psi))
for hns in (4, 8, 16, 32, 64, 128):
    h = hns * 1e-3
    ref = fine(t0, t0 + h, 64) @ psi
    big = cf4(t0, t0 + h) @ psi
    half = cf4(t0 + h / 2, t0 + h) @ (cf4(t0, t0 + h / 2) @ psi)
    rich = half + (half - big) / 15.0
    e4 = np.linalg.norm(big - ref); e4h = np.linalg.norm(half - ref); er = np.linalg.norm(rich - ref)
    print(f"h={hns:4d} ns  CF4 {e4:.2e} ({e4/h:.2e}/us)  2xCF4(h/2) {e4h:.2e}  Richardson {er:.2e} ({er/h:.2e}/us)")

# ---- CF6:5 from experiments/derive_cf.py (rows act first -> last) against Richardson(CF4)
F65 = np.array([[0.3196541055316575, -0.2695321497695407, 0.1692882382453091],
                [-0.0947598412780145, 0.0366853991682727, -0.0034216265944623],
                [0.5502114714927142, 0.0, -0.3317332233016935],
                [-0.0947598412780145, -0.0366853991682727, -0.0034216265944623],
                [0.3196541055316575, 0.2695321497695407, 0.1692882382453091]])
def cf65(a, b, v):
    h = b - a; mid = 0.5 * (a + b)
    a1 = h * Hm(mid); a2 = (h * h / 2.0) * Hp   # a3 = 0 for a linear sweep
    for row in F65:
        v = expmH(row[0] * a1 + row[1] * a2) @ v
    return v
print("CF6:5 vs Richardson(CF4), local error per step")
for hns in (8, 16, 24, 32, 48, 64):
    h = hns * 1e-3
    ref = fine(t0, t0 + h, 64) @ psi
    big = cf4(t0, t0 + h) @ psi
    half = cf4(t0 + h / 2, t0 + h) @ (cf4(t0, t0 + h / 2) @ psi)
    rich = half + (half - big) / 15.0
    c6 = cf65(t0, t0 + h, psi)
    print(f"h={hns:4d} ns  Richardson {np.linalg.norm(rich-ref):.2e}   CF6:5 {np.linalg.norm(c6-ref):.2e}")

# ---- single-exponential 4th-order Magnus exp(-i(B0 + i[B0,B1])) (exact commutator; Pulser-shaped when only the detuning varies)
def mag4(a, b):
    h = b - a; mid = 0.5 * (a + b)
    B0 = h * Hm(mid); B1 = (h * h / 12.0) * Hp
    return expmH(B0 + 1j * (B0 @ B1 - B1 @ B0))
print("Richardson on single-exponential Magnus-4 vs on CF4")
for hns in (8, 16, 24, 32, 48):
    h = hns * 1e-3
    ref = fine(t0, t0 + h, 64) @ psi
    out = {}
    for name, stepper in (("CF4", cf4), ("Mag4", mag4)):
        big = stepper(t0, t0 + h) @ psi
        half = stepper(t0 + h / 2, t0 + h) @ (stepper(t0, t0 + h / 2) @ psi)
        rich = half + (half - big) / 15.0
        out[name] = (np.linalg.norm(big - ref), np.linalg.norm(rich - ref))
    print(f"h={hns:4d} ns  CF4 {out['CF4'][0]:.2e} -> R {out['CF4'][1]:.2e}   Mag4 {out['Mag4'][0]:.2e} -> R {out['Mag4'][1]:.2e}")
