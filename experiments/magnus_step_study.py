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
