"""oracle/z2_trg_reference.py -- TEST INFRASTRUCTURE ONLY.
BASELINE configs[0]: runs the reference's own Z2 Ising TRG example (oracle/_ref/z2_trg, built by
`make -C oracle trg` from /root/reference/examples/z2_ising_trg.cpp) at bond dimension 32 on the host, sums the time the
example attributes to Contract and to SVD, and compares the free energy per site with Onsager's exact solution.

    python oracle/z2_trg_reference.py [chi=32] [threads=4]  > profiles/r1_config1_z2_trg_cpu.txt
"""
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def onsager_free_energy(beta: float, J: float = 1.0, n: int = 4096) -> float:
    """f = -(1/beta) [ ln 2 + 1/(2 pi^2) int_0^pi int_0^pi ln( cosh^2(2K) - sinh(2K) (cos a + cos b) ) da db / ... ]
    evaluated with the midpoint rule (the integrand is smooth away from the critical point)."""
    K = beta * J
    a = (np.arange(n) + 0.5) * np.pi / n
    ca = np.cos(a)
    g = np.log(np.cosh(2 * K) ** 2 - np.sinh(2 * K) * (ca[:, None] + ca[None, :]))
    return float(-(np.log(2.0) + 0.5 * g.mean()) / beta)


def main():
    chi = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    exe = os.path.join(HERE, "_ref", "z2_trg")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HERE, "trg"])
    env = dict(os.environ, OMP_WAIT_POLICY="passive")
    out = subprocess.run([exe, str(chi), str(threads), "1"], capture_output=True, text=True, env=env, check=True).stdout
    print(f"# reference examples/z2_ising_trg.cpp on the host, chi = {chi}, {threads} threads (HPTT + OpenBLAS 0.3.15)")
    print("# beta  f_trg  f_exact(Onsager)  |diff|  iterations  contract_s  svd_s  wall_s")
    contract = svd = 0.0
    for line in out.splitlines():
        m = re.search(r"SVD = ([\d.eE+-]+)s, Contract = ([\d.eE+-]+)s", line)
        if m:
            svd += float(m.group(1)); contract += float(m.group(2))
            continue
        m = re.match(r"^([\d.]+) (-?[\d.]+) (\d+) (\d+) ([\d.]+)$", line.strip())
        if m:
            beta, f = float(m.group(1)), float(m.group(2))
            fe = onsager_free_energy(beta)
            print(f"{beta:4.2f}  {f:.12f}  {fe:.12f}  {abs(f - fe):.2e}  {m.group(3)}  {contract:.3f}  {svd:.3f}  {m.group(5)}")
            contract = svd = 0.0


if __name__ == "__main__":
    main()
