"""CPU check for DESIGN.md section 10 item 3: can the IEEE division a = dot / nrm inside the 32 sequential decisions be
replaced by a multiplication with a precomputed reciprocal plus one FMA residual correction (Markstein) without
changing a single bit?   q0 = RN(x * r),  e = fma(-n, q0, x) (exact),  q1 = RN(q0 + e * r)  with r = RN(1 / n).
Exact rational arithmetic decides the correctly rounded quotient; fp32 roundings are emulated with numpy."""
import sys
from fractions import Fraction

import numpy as np


def rn32(fr):
    """Correctly rounded (nearest-even) float32 of a Fraction."""
    if fr == 0:
        return np.float32(0.0)
    f = np.float32(float(fr))          # double rounding can be off by one ulp: fix up against the neighbours
    cands = [f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))]
    best = min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(c.view(np.uint32)) & 1))
    return np.float32(best)


def markstein(x, n):
    r = rn32(Fraction(1) / Fraction(float(n)))
    q0 = rn32(Fraction(float(x)) * Fraction(float(r)))
    e = Fraction(float(x)) - Fraction(float(n)) * Fraction(float(q0))      # what fma(-n, q0, x) returns if exact
    e32 = rn32(e)
    exact_e = Fraction(float(e32)) == e
    q1 = rn32(Fraction(float(q0)) + Fraction(float(e32)) * Fraction(float(r)))
    return q1, exact_e


def main(samples):
    rng = np.random.default_rng(0)
    bad = inexact = 0
    for i in range(samples):
        if i % 4 == 0:      # adversarial divisors: significand close to all ones / exact powers of two
            n = np.float32(np.ldexp(2.0 - rng.integers(1, 64) * 2.0 ** -23, int(rng.integers(-3, 20))))
        else:
            n = np.float32(np.exp(rng.uniform(np.log(1e-2), np.log(1e6))))
        x = np.float32(rng.normal() * float(n) * np.exp(rng.uniform(-6, 2)))
        want = rn32(Fraction(float(x)) / Fraction(float(n)))
        got, exact_e = markstein(x, n)
        inexact += not exact_e
        if got.view(np.uint32) != want.view(np.uint32):
            bad += 1
            if bad <= 5:
                print("mismatch", float(x), float(n), float(got), float(want))
    print(f"{samples} samples: {bad} quotients differ from the correctly rounded one; residual not exactly "
          f"representable in fp32 in {inexact} cases")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 200000)
