"""TEST INFRASTRUCTURE (part of the oracle; never imported by the product path).

R's default random number stream, restated so the seeded inputs of the reference's own examples
(`set.seed(123); runif(...); rnorm(...)` in man/*.Rd, README.md and the vignette) can be regenerated
without R.  With them the oracle is pinned against the outputs the reference itself printed when its
documentation was rendered (docs/reference/*.html, vignettes/oem_vignette.html; see
tools/extract_reference_outputs.py and tests/test_reference_pins.py).

Third-party algorithm, not under /root/reference: GNU R (any version >= 3.0; the pieces below have not
changed since), files src/main/RNG.c and src/nmath/{snorm,runif,rnorm,qnorm}.c:

* `set.seed(s)` for the default kind "Mersenne-Twister": 50 rounds of the LCG `s = 69069 s + 1`
  (uint32), then 625 more rounds fill `dummy[0..624]`; `FixupSeeds` sets `dummy[0] = mti = 624`, so the
  first draw regenerates the block (RNG.c: RNG_Init, FixupSeeds).
* `unif_rand()` = MT19937 tempered word * 2.3283064365386963e-10, clamped into (0, 1) by `fixup`
  (RNG.c: MT_genrand, fixup).  numpy's MT19937 bit generator is the same recurrence and tempering, so
  its raw 32-bit outputs are reused once the key is set.
* `runif(n, a, b)` = a + (b - a) * unif_rand()            (runif.c)
* `norm_rand()` for the default normal kind "Inversion": `u = unif_rand(); u = (int)(2^27 u) +
  unif_rand(); qnorm5(u / 2^27)` (snorm.c) -- two uniforms per normal.
* `qnorm5` is Wichura's AS 241 (relative error ~1e-16).  It is evaluated here with
  `scipy.special.ndtri` (Cephes, same accuracy): the normals agree with R's to about one ulp, far
  below the 6-7 significant digits the reference printed.  Known answers from any R session:
  set.seed(123); runif(3) = 0.2875775 0.7883051 0.4089769; rnorm(3) = -0.56047565 -0.23017749
  1.55870831 (tests/test_reference_pins.py checks both).
"""
import numpy as np

_I2_32M1 = 2.328306437080797e-10      # RNG.c: fixup
_BIG = 134217728.0                    # snorm.c: BIG = 2^27


class RStream:
    def __init__(self, seed):
        s = np.uint32(seed & 0xFFFFFFFF)
        a, one = np.uint32(69069), np.uint32(1)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = a * s + one
            dummy = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = a * s + one
                dummy[j] = s
        self._bg = np.random.MT19937()
        self._bg.state = {"bit_generator": "MT19937", "state": {"key": dummy[1:].copy(), "pos": 624}}

    def unif_rand(self, n):
        u = self._bg.random_raw(int(n)).astype(np.float64) * 2.3283064365386963e-10
        u[u <= 0.0] = 0.5 * _I2_32M1
        u[(1.0 - u) <= 0.0] = 1.0 - 0.5 * _I2_32M1
        return u

    def runif(self, n, a=0.0, b=1.0):
        return a + (b - a) * self.unif_rand(n)

    def rnorm(self, n, mean=0.0, sd=1.0):
        from scipy.special import ndtri
        u = self.unif_rand(2 * int(n))
        v = (np.floor(_BIG * u[0::2]) + u[1::2]) / _BIG
        return mean + sd * ndtri(v)

    def matrix_rnorm(self, nrow, ncol, mean=0.0, sd=1.0):
        """matrix(rnorm(nrow * ncol, ...), nrow, ncol): column-major fill."""
        return np.asfortranarray(self.rnorm(nrow * ncol, mean, sd).reshape((ncol, nrow)).T)
