"""CPU ORACLE SUPPORT (test infrastructure, NOT product code) -- a minimal RHF + CCSD for water/STO-3G.

Purpose: pin the (T) oracle on a known answer the reference itself printed.  examples/Juliacon2022.ipynb:461-476 runs
`@energy ccsd(t)` for water / STO-3G / `df false` and prints (lines 497-615)
    Nuclear repulsion 8.8880641743, CCSD correlation -0.0537066985, CCSD energy -75.0187095932,
    Final (T) contribution -0.0000738086, CCSD(T) energy -75.0187834019.
This script rebuilds the inputs of `RCCSDpT(ccsd, moints, alg)` for that molecule from scratch -- McMurchie-Davidson
integrals over the STO-3G s/p Gaussians, RHF, spin-orbital CCSD (Stanton, Gauss, Watts, Bartlett, JCP 94, 4334 (1991)) --
and writes them, in the reference's array layouts, to tests/golden/water_sto3g.npz together with the energies it found.
tests/test_oracle_kat.py then feeds the stored T1/T2/integrals to the oracle and compares E(T) with the reference's print.

A second case pins it on a value the reference's own test suite asserts: water / 6-31G / df false, `@energy ccsd(t)` total
-76.121147867765558 (test/test_pT.jl:69-72, rtol 2e-8) -> tests/golden/water_631g.npz.

A third case is BASELINE config C2 itself: water / cc-pVTZ (o = 5, v = 53; d and f shells), for which test/test_pT.jl:5,31 holds
Psi4's CCSD(T) and CCSD totals, i.e. E(T) = -0.008051775570 -> tests/golden/water_ccpvtz.npz (integrals: oracle/mini_ints.py, numba).

A fourth case takes another molecule of the reference's test table: glycine / STO-3G (o = 20, v = 10; geometry test/xyz/glycine.xyz), for which
test/test_pT.jl:10,36 holds CCSD(T) and CCSD totals, E(T) = -0.007503098657 -> tests/golden/glycine_sto3g.npz.

A further case, ammonia / aug-cc-pVDZ (o = 5, v = 45, diffuse functions; test/test_pT.jl:6,32: E(T) = -0.005496117261), is kept like the
cc-pVTZ one -> tests/golden/ammonia_augccpvdz.npz.  And benzene / 6-31G (o = 21, v = 45; test/test_pT.jl:7,33: E(T) = -0.021110868073), is too big to keep as arrays (33 MB): only its
record is stored, tests/golden/pin_benzene_631g.json (oracle E(T) 3e-11 Eh from the held value; about 20 minutes on 8 cores).
Formaldehyde / 6-31G* (o = 8, v = 24 with spherical d shells; test/test_pT.jl:9,35: E(T) = -0.009118186572) -> tests/golden/formaldehyde_631gs.npz
(CCSD total 2e-12 Eh, oracle E(T) 6e-12 Eh from the held values; half a minute).
Likewise methane / cc-pVTZ (o = 5, v = 81; test/test_pT.jl:11,37: E(T) = -0.006426288342): tests/golden/pin_methane_ccpvtz.json (CCSD total
1e-11 Eh, oracle E(T) 5e-12 Eh from the held values; about 25 minutes, 37 GB of memory).
And ethanol / cc-pVDZ (o = 13, v = 59; test/test_pT.jl:8,34: E(T) = -0.012492525191): tests/golden/pin_ethanol_ccpvdz.json (CCSD total 8e-11 Eh,
oracle E(T) 1.2e-11 Eh from the held values; about 8 minutes).

Run from the repo root (pure-Python integrals: about a minute for sto-3g, several for 6-31g; numba: a few minutes for cc-pvtz):
    python oracle/mini_ccsd.py [sto-3g|6-31g|cc-pvtz|glycine/sto-3g|ammonia/aug-cc-pvdz|benzene/6-31g|methane/cc-pvtz|formaldehyde/6-31g*|ethanol/cc-pvdz] [numba]
"""
from __future__ import annotations

import itertools
import math
import os
import sys

import numpy as np
from scipy.special import hyp1f1

BOHR_TO_ANGSTROM = 0.529177210903  # src/Backend/PhysicalConstants.jl:38

# geometry of examples/Juliacon2022.ipynb:467-471 (Angstrom)
GEOM = [("O", (1.2091536548, 1.7664118189, -0.0171613972)),
        ("H", (2.1984800075, 1.7977100627, 0.0121161719)),
        ("H", (0.9197881882, 2.4580185570, 0.6297938830))]
Z = {"H": 1, "C": 6, "N": 7, "O": 8}
# other molecules of the reference's test suite: geometries as in /root/reference/test/xyz/<name>.xyz (Angstrom)
MOLECULES = {
    "water": GEOM,
    "glycine": [("C", (0.0000000, 0.5506150, 0.0000000)), ("O", (1.1776070, 0.8118730, 0.0000000)), ("O", (-0.9694240, 1.4927290, 0.0000000)),
                ("C", (-0.5847620, -0.8563810, 0.0000000)), ("N", (0.4006580, -1.9277700, 0.0000000)), ("H", (-0.5071000, 2.3458330, 0.0000000)),
                ("H", (-1.2456590, -0.9456040, 0.8813150)), ("H", (-1.2456590, -0.9456040, -0.8813150)),
                ("H", (1.0184570, -1.7812290, 0.8032340)), ("H", (1.0184570, -1.7812290, -0.8032340))],
    "ethanol": [("C", (1.1615830, -0.4067550, 0.0)), ("C", (0.0, 0.5527180, 0.0)), ("O", (-1.1871140, -0.2128600, 0.0)),
                ("H", (-1.9324340, 0.3838170, 0.0)), ("H", (2.1028600, 0.1358400, 0.0)), ("H", (1.1223470, -1.0398290, 0.8811340)),
                ("H", (1.1223470, -1.0398290, -0.8811340)), ("H", (0.0561470, 1.1935530, 0.8808960)), ("H", (0.0561470, 1.1935530, -0.8808960))],
    "formaldehyde": [("O", (0.0, 0.0, 0.6744930)), ("C", (0.0, 0.0, -0.5297240)), ("H", (0.0, 0.9347280, -1.1087990)),
                     ("H", (0.0, -0.9347280, -1.1087990))],
    "methane": [("C", (0.0, 0.0, 0.0)), ("H", (0.6268910, 0.6268910, 0.6268910)), ("H", (-0.6268910, -0.6268910, 0.6268910)),
                ("H", (-0.6268910, 0.6268910, -0.6268910)), ("H", (0.6268910, -0.6268910, -0.6268910))],
    "ammonia": [("N", (0.0, 0.0, 0.1173470)), ("H", (0.0, 0.9326490, -0.2738090)), ("H", (0.8076980, -0.4663250, -0.2738090)),
                ("H", (-0.8076980, -0.4663250, -0.2738090))],
    "benzene": [("C", (0.0000000, 1.3916730, 0.0)), ("C", (1.2052240, 0.6958360, 0.0)), ("C", (1.2052240, -0.6958360, 0.0)),
                ("C", (0.0000000, -1.3916730, 0.0)), ("C", (-1.2052240, -0.6958360, 0.0)), ("C", (-1.2052240, 0.6958360, 0.0)),
                ("H", (0.0000000, 2.4695880, 0.0)), ("H", (2.1387260, 1.2347940, 0.0)), ("H", (2.1387260, -1.2347940, 0.0)),
                ("H", (0.0000000, -2.4695880, 0.0)), ("H", (-2.1387260, -1.2347940, 0.0)), ("H", (-2.1387260, 1.2347940, 0.0))],
}

# STO-3G (Basis Set Exchange), shells as (l, exponents, coefficients)
STO3G = {
    "H": [(0, [0.3425250914e+01, 0.6239137298e+00, 0.1688554040e+00], [0.1543289673e+00, 0.5353281423e+00, 0.4446345422e+00])],
    "O": [(0, [0.1307093214e+03, 0.2380886605e+02, 0.6443608313e+01], [0.1543289673e+00, 0.5353281423e+00, 0.4446345422e+00]),
          (0, [0.5033151319e+01, 0.1169596125e+01, 0.3803889600e+00], [-0.9996722919e-01, 0.3995128261e+00, 0.7001154689e+00]),
          (1, [0.5033151319e+01, 0.1169596125e+01, 0.3803889600e+00], [0.1559162750e+00, 0.6076837186e+00, 0.3919573931e+00])],
    "C": [(0, [0.7161683735e+02, 0.1304509632e+02, 0.3530512160e+01], [0.1543289673e+00, 0.5353281423e+00, 0.4446345422e+00]),
          (0, [0.2941249355e+01, 0.6834830964e+00, 0.2222899159e+00], [-0.9996722919e-01, 0.3995128261e+00, 0.7001154689e+00]),
          (1, [0.2941249355e+01, 0.6834830964e+00, 0.2222899159e+00], [0.1559162750e+00, 0.6076837186e+00, 0.3919573931e+00])],
    "N": [(0, [0.9910616896e+02, 0.1805231239e+02, 0.4885660238e+01], [0.1543289673e+00, 0.5353281423e+00, 0.4446345422e+00]),
          (0, [0.3780455879e+01, 0.8784966449e+00, 0.2857143744e+00], [-0.9996722919e-01, 0.3995128261e+00, 0.7001154689e+00]),
          (1, [0.3780455879e+01, 0.8784966449e+00, 0.2857143744e+00], [0.1559162750e+00, 0.6076837186e+00, 0.3919573931e+00])],
}


# 6-31G (Basis Set Exchange; Hehre, Ditchfield, Pople JCP 56, 2257 (1972)); the SP shells are written out as an s and a p shell
B631G = {
    "H": [(0, [18.7311370, 2.8253937, 0.6401217], [0.03349460, 0.23472695, 0.81375733]),
          (0, [0.1612778], [1.0])],
    "O": [(0, [5484.6717000, 825.2349500, 188.0469600, 52.9645000, 16.8975700, 5.7996353],
              [0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.4701930, 0.3585209]),
          (0, [15.5396160, 3.5999336, 1.0137618], [-0.1107775, -0.1480263, 1.1307670]),
          (1, [15.5396160, 3.5999336, 1.0137618], [0.0708743, 0.3397528, 0.7271586]),
          (0, [0.2700058], [1.0]),
          (1, [0.2700058], [1.0])],
    "C": [(0, [3047.5249000, 457.3695100, 103.9486900, 29.2101550, 9.2866630, 3.1639270],
              [0.0018347, 0.0140373, 0.0688426, 0.2321844, 0.4679413, 0.3623120]),
          (0, [7.8682724, 1.8812885, 0.5442493], [-0.1193324, -0.1608542, 1.1434564]),
          (1, [7.8682724, 1.8812885, 0.5442493], [0.0689991, 0.3164240, 0.7443083]),
          (0, [0.1687144], [1.0]),
          (1, [0.1687144], [1.0])],
}
# cc-pVTZ (Dunning, JCP 90, 1007 (1989); Basis Set Exchange, optimised general contractions): O (10s5p2d1f) -> [4s3p2d1f],
# H (5s2p1d) -> [3s2p1d]; real solid harmonics (58 functions for water).  Integrals: oracle/mini_ints.py.
CCPVTZ = {
    "H": [(0, [33.87, 5.095, 1.159], [0.006068, 0.045308, 0.202822]),
          (0, [0.3258], [1.0]), (0, [0.1027], [1.0]),
          (1, [1.407], [1.0]), (1, [0.388], [1.0]),
          (2, [1.057], [1.0])],
    "O": [(0, [15330.0, 2299.0, 522.4, 147.3, 47.55, 16.76, 6.207, 0.6882],
              [0.000508, 0.003929, 0.020243, 0.079181, 0.230687, 0.433118, 0.350260, -0.008154]),
          (0, [15330.0, 2299.0, 522.4, 147.3, 47.55, 16.76, 6.207, 0.6882],
              [-0.000115, -0.000895, -0.004636, -0.018724, -0.058463, -0.136463, -0.175740, 0.603418]),
          (0, [1.752], [1.0]), (0, [0.2384], [1.0]),
          (1, [34.46, 7.749, 2.280], [0.015928, 0.099740, 0.310492]),
          (1, [0.7156], [1.0]), (1, [0.2140], [1.0]),
          (2, [2.314], [1.0]), (2, [0.645], [1.0]),
          (3, [1.428], [1.0])],
    "C": [(0, [8236.0, 1235.0, 280.8, 79.27, 25.59, 8.997, 3.319, 0.3643],
              [0.000531, 0.004108, 0.021087, 0.081853, 0.234817, 0.434401, 0.346129, -0.008983]),
          (0, [8236.0, 1235.0, 280.8, 79.27, 25.59, 8.997, 3.319, 0.3643],
              [-0.000113, -0.000878, -0.004540, -0.018133, -0.055760, -0.126895, -0.170352, 0.598684]),
          (0, [0.9059], [1.0]), (0, [0.1285], [1.0]),
          (1, [18.71, 4.133, 1.200], [0.014031, 0.086866, 0.290216]),
          (1, [0.3827], [1.0]), (1, [0.1209], [1.0]),
          (2, [1.097], [1.0]), (2, [0.318], [1.0]),
          (3, [0.761], [1.0])],
}
# aug-cc-pVDZ (Dunning 1989; Kendall, Dunning, Harrison 1992): N (10s5p2d) -> [4s3p2d], H (5s2p) -> [3s2p]
AUGCCPVDZ = {
    "H": [(0, [13.01, 1.962, 0.4446], [0.019685, 0.137977, 0.478148]),
          (0, [0.122], [1.0]), (0, [0.02974], [1.0]),
          (1, [0.727], [1.0]), (1, [0.141], [1.0])],
    "N": [(0, [9046.0, 1357.0, 309.3, 87.73, 28.56, 10.21, 3.838, 0.7466],
              [0.000700, 0.005389, 0.027406, 0.103207, 0.278723, 0.448540, 0.278238, 0.015440]),
          (0, [9046.0, 1357.0, 309.3, 87.73, 28.56, 10.21, 3.838, 0.7466],
              [-0.000153, -0.001208, -0.005992, -0.024544, -0.067459, -0.158078, -0.121831, 0.549003]),
          (0, [0.2248], [1.0]), (0, [0.06124], [1.0]),
          (1, [13.55, 2.917, 0.7973], [0.039919, 0.217169, 0.510319]),
          (1, [0.2185], [1.0]), (1, [0.05611], [1.0]),
          (2, [0.817], [1.0]), (2, [0.230], [1.0])],
}
# cc-pVDZ (Dunning 1989): C, O (9s4p1d) -> [3s2p1d], H (4s1p) -> [2s1p]
CCPVDZ = {
    "H": [(0, [13.01, 1.962, 0.4446], [0.019685, 0.137977, 0.478148]),
          (0, [0.122], [1.0]),
          (1, [0.727], [1.0])],
    "C": [(0, [6665.0, 1000.0, 228.0, 64.71, 21.06, 7.495, 2.797, 0.5215],
              [0.000692, 0.005329, 0.027077, 0.101718, 0.274740, 0.448564, 0.285074, 0.015204]),
          (0, [6665.0, 1000.0, 228.0, 64.71, 21.06, 7.495, 2.797, 0.5215],
              [-0.000146, -0.001154, -0.005725, -0.023312, -0.063955, -0.149981, -0.127262, 0.544529]),
          (0, [0.1596], [1.0]),
          (1, [9.439, 2.002, 0.5456], [0.038109, 0.209480, 0.508557]),
          (1, [0.1517], [1.0]),
          (2, [0.55], [1.0])],
    "O": [(0, [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013],
              [0.000710, 0.005470, 0.027837, 0.104800, 0.283062, 0.448719, 0.270952, 0.015458]),
          (0, [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013],
              [-0.000160, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368]),
          (0, [0.3023], [1.0]),
          (1, [17.70, 3.854, 1.046], [0.043018, 0.228913, 0.508728]),
          (1, [0.2753], [1.0]),
          (2, [1.185], [1.0])],
}
# 6-31G* (Hariharan, Pople 1973): 6-31G plus one d shell (exponent 0.8) on C and O.  Five real solid harmonics per d shell, like every
# other set here: the values the reference holds for formaldehyde correspond to that (six Cartesian d functions give a CCSD total
# 6.9 mEh lower -- tried, does not match).
B631GS = {sym: shells + ([(2, [0.8], [1.0])] if sym != "H" else []) for sym, shells in B631G.items()}
BASES = {"sto-3g": STO3G, "6-31g": B631G, "cc-pvtz": CCPVTZ, "aug-cc-pvdz": AUGCCPVDZ, "6-31g*": B631GS, "cc-pvdz": CCPVDZ}
# What the reference holds for each case: the printed run of examples/Juliacon2022.ipynb:497-615 (STO-3G) and the Psi4 total
# energy its own test asserts for `@energy ccsd(t)`, water / 6-31G / df false (test/test_pT.jl:69-72, rtol 2e-8).
REFERENCE = {"sto-3g": {"e_nuc": 8.8880641743, "e_corr": -0.0537066985, "e_ccsd": -75.0187095932, "e_t": -0.0000738086,
                        "e_ccsd_t": -75.0187834019},
             "6-31g": {"e_ccsd_t": -76.121147867765558},
             # Psi4 totals the reference's test suite holds for water / cc-pVTZ / df false (test/test_pT.jl:5 Econv[1], :31 CCSDconv[1])
             "cc-pvtz": {"e_ccsd": -76.335767822597347, "e_ccsd_t": -76.343819598166903, "e_t": -76.343819598166903 + 76.335767822597347},
             # glycine / STO-3G / df false: test/test_pT.jl:10 Econv[6], :36 CCSDconv[6] (o = 20, v = 10)
             # benzene / 6-31G / df false: test/test_pT.jl:7 Econv[3], :33 CCSDconv[3] (o = 21, v = 45)
             "benzene/6-31g": {"e_ccsd": -231.188695053088594, "e_ccsd_t": -231.209805921161490, "e_t": -231.209805921161490 + 231.188695053088594},
             # ethanol / cc-pVDZ / df false: test/test_pT.jl:8 Econv[4], :34 CCSDconv[4] (o = 13, v = 59)
             "ethanol/cc-pvdz": {"e_ccsd": -154.616795317070142, "e_ccsd_t": -154.629287842261505, "e_t": -154.629287842261505 + 154.616795317070142},
             # formaldehyde / 6-31G* / df false: test/test_pT.jl:9 Econv[5], :35 CCSDconv[5] (o = 8, v = 24: spherical d shells)
             "formaldehyde/6-31g*": {"e_ccsd": -114.180708994251702, "e_ccsd_t": -114.189827180824139, "e_t": -114.189827180824139 + 114.180708994251702},
             # methane / cc-pVTZ / df false: test/test_pT.jl:11 Econv[7], :37 CCSDconv[7] (o = 5, v = 81)
             "methane/cc-pvtz": {"e_ccsd": -40.448675124014166, "e_ccsd_t": -40.455101412356250, "e_t": -40.455101412356250 + 40.448675124014166},
             # ammonia / aug-cc-pVDZ / df false: test/test_pT.jl:6 Econv[2], :32 CCSDconv[2] (o = 5, v = 45)
             "ammonia/aug-cc-pvdz": {"e_ccsd": -56.422272522003723, "e_ccsd_t": -56.427768639264869, "e_t": -56.427768639264869 + 56.422272522003723},
             "glycine/sto-3g": {"e_ccsd": -279.415437830677774, "e_ccsd_t": -279.422940929335255, "e_t": -279.422940929335255 + 279.415437830677774}}


def dfact(n):
    return 1 if n <= 0 else n * dfact(n - 2)


class BF:
    """contracted cartesian Gaussian"""

    def __init__(self, origin, lmn, exps, coefs):
        self.o = np.array(origin, float)
        self.lmn = lmn
        self.exps = list(exps)
        l, m, n = lmn
        self.norm = [(2 * a / math.pi) ** 0.75 * (4 * a) ** ((l + m + n) / 2) / math.sqrt(dfact(2 * l - 1) * dfact(2 * m - 1) * dfact(2 * n - 1))
                     for a in exps]
        self.coefs = list(coefs)
        # normalise the contraction
        s = 0.0
        for a, ca, na in zip(exps, coefs, self.norm):
            for b, cb, nb in zip(exps, coefs, self.norm):
                s += ca * cb * na * nb * overlap_prim(a, lmn, self.o, b, lmn, self.o)
        self.coefs = [c / math.sqrt(s) for c in coefs]


def E(i, j, t, Qx, a, b):
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return math.exp(-q * Qx * Qx)
    if j == 0:
        return (1 / (2 * p)) * E(i - 1, j, t - 1, Qx, a, b) - (q * Qx / a) * E(i - 1, j, t, Qx, a, b) + (t + 1) * E(i - 1, j, t + 1, Qx, a, b)
    return (1 / (2 * p)) * E(i, j - 1, t - 1, Qx, a, b) + (q * Qx / b) * E(i, j - 1, t, Qx, a, b) + (t + 1) * E(i, j - 1, t + 1, Qx, a, b)


def overlap_prim(a, lmn1, A, b, lmn2, B):
    S = 1.0
    for x in range(3):
        S *= E(lmn1[x], lmn2[x], 0, A[x] - B[x], a, b)
    return S * (math.pi / (a + b)) ** 1.5


def kinetic_prim(a, lmn1, A, b, lmn2, B):
    l2, m2, n2 = lmn2
    t0 = b * (2 * (l2 + m2 + n2) + 3) * overlap_prim(a, lmn1, A, b, lmn2, B)
    t1 = -2 * b * b * (overlap_prim(a, lmn1, A, b, (l2 + 2, m2, n2), B) + overlap_prim(a, lmn1, A, b, (l2, m2 + 2, n2), B)
                       + overlap_prim(a, lmn1, A, b, (l2, m2, n2 + 2), B))
    t2 = -0.5 * (l2 * (l2 - 1) * overlap_prim(a, lmn1, A, b, (l2 - 2, m2, n2), B) + m2 * (m2 - 1) * overlap_prim(a, lmn1, A, b, (l2, m2 - 2, n2), B)
                 + n2 * (n2 - 1) * overlap_prim(a, lmn1, A, b, (l2, m2, n2 - 2), B))
    return t0 + t1 + t2


def boys(n, x):
    return hyp1f1(n + 0.5, n + 1.5, -x) / (2 * n + 1)


def R(t, u, v, n, p, PC, RPC2):
    if t < 0 or u < 0 or v < 0:
        return 0.0
    if t == u == v == 0:
        return (-2 * p) ** n * boys(n, p * RPC2)
    if t == u == 0:
        val = PC[2] * R(t, u, v - 1, n + 1, p, PC, RPC2)
        if v > 1:
            val += (v - 1) * R(t, u, v - 2, n + 1, p, PC, RPC2)
        return val
    if t == 0:
        val = PC[1] * R(t, u - 1, v, n + 1, p, PC, RPC2)
        if u > 1:
            val += (u - 1) * R(t, u - 2, v, n + 1, p, PC, RPC2)
        return val
    val = PC[0] * R(t - 1, u, v, n + 1, p, PC, RPC2)
    if t > 1:
        val += (t - 1) * R(t - 2, u, v, n + 1, p, PC, RPC2)
    return val


def nuclear_prim(a, lmn1, A, b, lmn2, B, C):
    p = a + b
    P = (a * A + b * B) / p
    PC = P - C
    RPC2 = float(PC @ PC)
    val = 0.0
    for t in range(lmn1[0] + lmn2[0] + 1):
        for u in range(lmn1[1] + lmn2[1] + 1):
            for v in range(lmn1[2] + lmn2[2] + 1):
                val += (E(lmn1[0], lmn2[0], t, A[0] - B[0], a, b) * E(lmn1[1], lmn2[1], u, A[1] - B[1], a, b)
                        * E(lmn1[2], lmn2[2], v, A[2] - B[2], a, b) * R(t, u, v, 0, p, PC, RPC2))
    return 2 * math.pi / p * val


def eri_prim(a, lmn1, A, b, lmn2, B, c, lmn3, C, d, lmn4, D):
    p, q = a + b, c + d
    alpha = p * q / (p + q)
    P = (a * A + b * B) / p
    Q = (c * C + d * D) / q
    PQ = P - Q
    RPQ2 = float(PQ @ PQ)
    Eab = [[E(lmn1[x], lmn2[x], t, A[x] - B[x], a, b) for t in range(lmn1[x] + lmn2[x] + 1)] for x in range(3)]
    Ecd = [[E(lmn3[x], lmn4[x], t, C[x] - D[x], c, d) for t in range(lmn3[x] + lmn4[x] + 1)] for x in range(3)]
    val = 0.0
    for t, et in enumerate(Eab[0]):
        for u, eu in enumerate(Eab[1]):
            for v, ev in enumerate(Eab[2]):
                for tau, ft in enumerate(Ecd[0]):
                    for nu, fu in enumerate(Ecd[1]):
                        for phi, fv in enumerate(Ecd[2]):
                            val += et * eu * ev * ft * fu * fv * (-1) ** (tau + nu + phi) * R(t + tau, u + nu, v + phi, 0, alpha, PQ, RPQ2)
    return val * 2 * math.pi ** 2.5 / (p * q * math.sqrt(p + q))


def contracted(fn, *bfs, extra=()):
    val = 0.0
    for idx in itertools.product(*[range(len(b.exps)) for b in bfs]):
        c = 1.0
        args = []
        for b, i in zip(bfs, idx):
            c *= b.coefs[i] * b.norm[i]
            args += [b.exps[i], b.lmn, b.o]
        val += c * fn(*args, *extra)
    return val


def build_basis(basis="sto-3g"):
    bfs, atoms = [], []
    for sym, xyz in GEOM:
        pos = np.array(xyz) / BOHR_TO_ANGSTROM
        atoms.append((Z[sym], pos))
        for l, exps, coefs in BASES[basis][sym]:
            for lmn in ([(0, 0, 0)] if l == 0 else [(1, 0, 0), (0, 1, 0), (0, 0, 1)]):
                bfs.append(BF(pos, lmn, exps, coefs))
    return bfs, atoms


def integrals(bfs, atoms):
    n = len(bfs)
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            S[i, j] = S[j, i] = contracted(overlap_prim, bfs[i], bfs[j])
            T[i, j] = T[j, i] = contracted(kinetic_prim, bfs[i], bfs[j])
            v = 0.0
            for z, pos in atoms:
                v -= z * contracted(nuclear_prim, bfs[i], bfs[j], extra=(pos,))
            V[i, j] = V[j, i] = v
    ERI = np.zeros((n, n, n, n))
    for i in range(n):
        for j in range(i + 1):
            for k in range(n):
                for l in range(k + 1):
                    if i * (i + 1) // 2 + j < k * (k + 1) // 2 + l:
                        continue
                    x = contracted(eri_prim, bfs[i], bfs[j], bfs[k], bfs[l])
                    for (a, b, c, d) in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k), (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                        ERI[a, b, c, d] = x
    enuc = sum(atoms[a][0] * atoms[b][0] / np.linalg.norm(atoms[a][1] - atoms[b][1]) for a in range(len(atoms)) for b in range(a))
    return S, T, V, ERI, enuc


def rhf(S, H, ERI, ndocc, tol=1e-13, maxit=500):
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T
    F = H.copy()
    D = np.zeros_like(H)
    e_old = 0.0
    fs, es = [], []
    for it in range(maxit):
        Fp = X @ F @ X
        eps, C2 = np.linalg.eigh(Fp)
        C = X @ C2
        Cocc = C[:, :ndocc]
        D = Cocc @ Cocc.T
        J = np.einsum("pqrs,rs->pq", ERI, D)
        K = np.einsum("prqs,rs->pq", ERI, D)
        F = H + 2 * J - K
        e = float(np.sum(D * (H + F)))
        # DIIS
        err = X @ (F @ D @ S - S @ D @ F) @ X
        fs.append(F.copy()); es.append(err)
        fs, es = fs[-8:], es[-8:]
        if len(fs) > 1:
            nB = len(fs)
            B = -np.ones((nB + 1, nB + 1)); B[-1, -1] = 0
            for a in range(nB):
                for b in range(nB):
                    B[a, b] = np.sum(es[a] * es[b])
            rhs = np.zeros(nB + 1); rhs[-1] = -1
            try:
                c = np.linalg.solve(B, rhs)[:-1]
                F = sum(ci * fi for ci, fi in zip(c, fs))
            except np.linalg.LinAlgError:
                pass
        if abs(e - e_old) < tol and np.max(np.abs(err)) < 1e-11:
            break
        e_old = e
    Fp = X @ (H + 2 * np.einsum("pqrs,rs->pq", ERI, D) - np.einsum("prqs,rs->pq", ERI, D)) @ X
    eps, C2 = np.linalg.eigh(Fp)
    C = X @ C2
    Cocc = C[:, :ndocc]
    D = Cocc @ Cocc.T
    Ffin = H + 2 * np.einsum("pqrs,rs->pq", ERI, D) - np.einsum("prqs,rs->pq", ERI, D)
    return float(np.sum(D * (H + Ffin))), eps, C


def ccsd_spinorbital(eps, MO, ndocc, tol=1e-13, maxit=200):
    """MO: (pq|rs) chemist, spatial.  Returns correlation energy, t1, t2 (spin orbital, interleaved alpha/beta)."""
    n = len(eps)
    ns = 2 * n
    sp = np.arange(ns) // 2
    spin = np.arange(ns) % 2
    g = MO[np.ix_(sp, sp, sp, sp)] * (spin[:, None, None, None] == spin[None, :, None, None]) * (spin[None, None, :, None] == spin[None, None, None, :])
    phys = g.transpose(0, 2, 1, 3)          # <pq|rs> = (pr|qs)
    A = phys - phys.transpose(0, 1, 3, 2)   # <pq||rs>
    fs = np.repeat(eps, 2)
    no = 2 * ndocc
    o, v = slice(0, no), slice(no, ns)
    fo, fv = fs[o], fs[v]
    Dia = fo[:, None] - fv[None, :]
    Dijab = fo[:, None, None, None] + fo[None, :, None, None] - fv[None, None, :, None] - fv[None, None, None, :]
    t1 = np.zeros((no, ns - no))
    t2 = A[o, o, v, v] / Dijab
    es = lambda *a: np.einsum(*a, optimize=True)
    e_old = 0.0
    hist_t, hist_e = [], []
    for it in range(maxit):
        tau_t = t2 + 0.5 * (es("ia,jb->ijab", t1, t1) - es("ib,ja->ijab", t1, t1))
        tau = t2 + es("ia,jb->ijab", t1, t1) - es("ib,ja->ijab", t1, t1)
        Fae = es("mf,mafe->ae", t1, A[o, v, v, v]) - 0.5 * es("mnaf,mnef->ae", tau_t, A[o, o, v, v])
        Fmi = es("ne,mnie->mi", t1, A[o, o, o, v]) + 0.5 * es("inef,mnef->mi", tau_t, A[o, o, v, v])
        Fme = es("nf,mnef->me", t1, A[o, o, v, v])
        Wmnij = A[o, o, o, o] + es("je,mnie->mnij", t1, A[o, o, o, v]) - es("ie,mnje->mnij", t1, A[o, o, o, v]) + 0.25 * es("ijef,mnef->mnij", tau, A[o, o, v, v])
        Wabef = A[v, v, v, v] - es("mb,amef->abef", t1, A[v, o, v, v]) + es("ma,bmef->abef", t1, A[v, o, v, v]) + 0.25 * es("mnab,mnef->abef", tau, A[o, o, v, v])
        Wmbej = (A[o, v, v, o] + es("jf,mbef->mbej", t1, A[o, v, v, v]) - es("nb,mnej->mbej", t1, A[o, o, v, o])
                 - es("jnfb,mnef->mbej", 0.5 * t2 + es("jf,nb->jnfb", t1, t1), A[o, o, v, v]))
        r1 = (es("ie,ae->ia", t1, Fae) - es("ma,mi->ia", t1, Fmi) + es("imae,me->ia", t2, Fme) - es("nf,naif->ia", t1, A[o, v, o, v])
              - 0.5 * es("imef,maef->ia", t2, A[o, v, v, v]) - 0.5 * es("mnae,nmei->ia", t2, A[o, o, v, o]))
        r2 = A[o, o, v, v].copy()
        tmp = es("ijae,be->ijab", t2, Fae - 0.5 * es("mb,me->be", t1, Fme))
        r2 += tmp - tmp.transpose(0, 1, 3, 2)
        tmp = es("imab,mj->ijab", t2, Fmi + 0.5 * es("je,me->mj", t1, Fme))
        r2 -= tmp - tmp.transpose(1, 0, 2, 3)
        r2 += 0.5 * es("mnab,mnij->ijab", tau, Wmnij) + 0.5 * es("ijef,abef->ijab", tau, Wabef)
        tmp = es("imae,mbej->ijab", t2, Wmbej) - es("ie,ma,mbej->ijab", t1, t1, A[o, v, v, o])
        r2 += tmp - tmp.transpose(1, 0, 2, 3) - tmp.transpose(0, 1, 3, 2) + tmp.transpose(1, 0, 3, 2)
        tmp = es("ie,abej->ijab", t1, A[v, v, v, o])
        r2 += tmp - tmp.transpose(1, 0, 2, 3)
        tmp = es("ma,mbij->ijab", t1, A[o, v, o, o])
        r2 -= tmp - tmp.transpose(0, 1, 3, 2)
        t1n, t2n = r1 / Dia, r2 / Dijab
        # DIIS on the amplitudes
        vec = np.concatenate([t1n.ravel(), t2n.ravel()])
        err = vec - np.concatenate([t1.ravel(), t2.ravel()])
        hist_t.append(vec); hist_e.append(err)
        hist_t, hist_e = hist_t[-8:], hist_e[-8:]
        if len(hist_t) > 2:
            nB = len(hist_t)
            B = -np.ones((nB + 1, nB + 1)); B[-1, -1] = 0
            for a in range(nB):
                for b in range(nB):
                    B[a, b] = hist_e[a] @ hist_e[b]
            rhs = np.zeros(nB + 1); rhs[-1] = -1
            try:
                c = np.linalg.solve(B, rhs)[:-1]
                vec = sum(ci * ti for ci, ti in zip(c, hist_t))
            except np.linalg.LinAlgError:
                pass
        t1 = vec[:t1.size].reshape(t1.shape)
        t2 = vec[t1.size:].reshape(t2.shape)
        e = 0.25 * es("ijab,ijab->", A[o, o, v, v], t2) + 0.5 * es("ijab,ia,jb->", A[o, o, v, v], t1, t1)
        if abs(e - e_old) < tol and np.max(np.abs(err)) < 1e-12:
            break
        e_old = e
    return float(e), t1, t2, A, fs, no


def pt_spinorbital(t1, t2, A, fs, no):
    """Textbook spin-orbital (T) -- an independent check of the closed-shell oracle on the molecular amplitudes."""
    ns = len(fs)
    o, v = slice(0, no), slice(no, ns)
    es = np.einsum
    fo, fv = fs[o], fs[v]
    D = (fo[:, None, None, None, None, None] + fo[None, :, None, None, None, None] + fo[None, None, :, None, None, None]
         - fv[None, None, None, :, None, None] - fv[None, None, None, None, :, None] - fv[None, None, None, None, None, :])

    def perm(x):
        y = x - x.transpose(1, 0, 2, 3, 4, 5) - x.transpose(2, 1, 0, 3, 4, 5)
        return y - y.transpose(0, 1, 2, 4, 3, 5) - y.transpose(0, 1, 2, 5, 4, 3)

    conn = perm(es("jkae,eibc->ijkabc", t2, A[v, o, v, v]) - es("imbc,majk->ijkabc", t2, A[o, v, o, o]))
    disc = perm(es("ia,jkbc->ijkabc", t1, A[o, o, v, v]))
    return float(np.sum(conn * (conn + disc) / D) / 36.0)


def integrals_numba(basis):
    """The same integrals from oracle/mini_ints.py (numba, any angular momentum, real solid harmonics)."""
    from oracle import mini_ints
    shells, atoms = [], []
    for sym, xyz in GEOM:
        pos = np.array(xyz) / BOHR_TO_ANGSTROM
        atoms.append((Z[sym], pos))
        for l, exps, coefs in BASES[basis][sym]:
            shells.append((pos, l, exps, coefs))
    return mini_ints.integrals(shells, atoms)


def main(basis="sto-3g", engine="auto"):
    global GEOM
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    molecule, case = "water", basis
    if "/" in basis:                      # "glycine/sto-3g"
        molecule, basis = basis.split("/")
        engine = "numba" if engine == "auto" else engine
    GEOM = MOLECULES[molecule]
    if engine == "numba" or (engine == "auto" and basis == "cc-pvtz"):
        S, T, V, ERI, enuc = integrals_numba(basis)
    else:
        bfs, atoms = build_basis(basis)
        S, T, V, ERI, enuc = integrals(bfs, atoms)
    ndocc = sum(Z[sym] for sym, _ in GEOM) // 2
    e_el, eps, C = rhf(S, T + V, ERI, ndocc)
    e_rhf = e_el + enuc
    MO = np.einsum("pqrs,pi,qj,rk,sl->ijkl", ERI, C, C, C, C, optimize=True)
    e_cc, t1, t2, A, fs, no = ccsd_spinorbital(eps, MO, ndocc)
    # the 6-index spin-orbital (T) is an independent check for the small cases only ((2o)^3 (2v)^3 doubles)
    e_t_so = pt_spinorbital(t1, t2, A, fs, no) if len(eps) <= 16 else float("nan")
    o, n = ndocc, len(eps)
    v = n - o
    T1 = np.asfortranarray(t1[0::2, 0::2])
    T2 = np.asfortranarray(t2[0::2, 1::2, 0::2, 1::2])
    oc, vi = slice(0, o), slice(o, n)
    OVVV = np.asfortranarray(MO[oc, vi, vi, vi])
    OOOV = np.asfortranarray(MO[oc, oc, oc, vi])
    OVOV = np.asfortranarray(MO[oc, vi, oc, vi])
    fo, fv = eps[:o].copy(), eps[o:].copy()
    from oracle import pt_numpy as P
    e_t = P.pt_ijk(T1, T2, OVVV, OOOV, OVOV, fo, fv)
    ref = REFERENCE[case]
    note = lambda k: f"   (reference {ref[k]:.10f})" if k in ref else ""
    print(f"{molecule} / {basis}: o={o} v={v}")
    print(f"E_nuc   {enuc:.10f}" + note("e_nuc"))
    print(f"E_RHF   {e_rhf:.10f}")
    print(f"E_corr  {e_cc:.10f}" + note("e_corr"))
    print(f"E_CCSD  {e_rhf + e_cc:.10f}" + note("e_ccsd"))
    print(f"E(T)    {e_t:.10f}" + note("e_t") + f"   spin-orbital formula: {e_t_so:.10f}")
    print(f"CCSD(T) {e_rhf + e_cc + e_t:.10f}" + note("e_ccsd_t"))
    out = os.path.join(root, "tests", "golden", molecule + "_" + basis.replace("-", "").replace("*", "s") + ".npz")
    extra = {}
    if o * v ** 3 * 8 > (8 << 20):   # (ia|bc) alone above 8 MB (benzene / 6-31G: 33 MB in all): keep the record, not the arrays
        import json
        rec = {"molecule": molecule, "basis": basis, "o": o, "v": v, "e_nuc": enuc, "e_rhf": e_rhf, "e_corr": e_cc, "e_ccsd": e_rhf + e_cc,
               "e_t_oracle_pt_ijk": e_t, "e_ccsd_t": e_rhf + e_cc + e_t, "reference": ref,
               "d_e_ccsd": e_rhf + e_cc - ref.get("e_ccsd", float("nan")), "d_e_t": e_t - ref.get("e_t", float("nan")),
               "command": "python oracle/mini_ccsd.py " + case}
        path = os.path.join(root, "tests", "golden", "pin_" + molecule + "_" + basis.replace("-", "").replace("*", "s") + ".json")
        json.dump(rec, open(path, "w"), indent=1)
        print("wrote", path)
        return
    if o * v ** 3 * 8 > (1 << 20):   # e.g. 5 x 53^3 doubles: keep only b >= c of (ia|bc) = (ia|cb) (tests/test_oracle_kat.py unpacks it)
        iu = np.triu_indices(v)
        packed = np.ascontiguousarray(OVVV[:, :, iu[1], iu[0]])      # [i, a, (b >= c)]
        assert np.max(np.abs(OVVV - OVVV.transpose(0, 1, 3, 2))) < 1e-12
        np.savez_compressed(out, T1=T1, T2=T2, OVVV_packed=packed, OOOV=OOOV, OVOV=OVOV, fo=fo, fv=fv, e_nuc=enuc, e_rhf=e_rhf, e_corr=e_cc, e_t=e_t)
        print("wrote", out)
        return
    if basis == "sto-3g" and molecule == "water":   # small enough to keep: lets the AO -> MO route (fpt_triples_ao) be checked on a real molecule
        extra = {"AOERI": np.asfortranarray(ERI), "C": np.asfortranarray(C)}
    np.savez(out, T1=T1, T2=T2, OVVV=OVVV, OOOV=OOOV, OVOV=OVOV, fo=fo, fv=fv, e_nuc=enuc, e_rhf=e_rhf, e_corr=e_cc, e_t=e_t,
             e_t_spinorbital=e_t_so, **extra)
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "sto-3g", sys.argv[2] if len(sys.argv) > 2 else "auto")
