"""CPU oracle of the integrator -- TEST INFRASTRUCTURE ONLY.

Numpy float64 restatement of one outer step of BasicIntegrator::dynamic (src/md/integrator.cpp:70-170) for the
velocity-Verlet (nrespa = 1) and r-RESPA propagators (RespaDevice::velR0/R1/R2, src/md/propagator.cpp:170-187;
mdPos / mdVel / mdVel2, src/acc/mdpq.cpp:12-115), the kinetic energy / temperature of kinetic() and the Bussi
velocity-rescale formula of bussiThermostat (src/mdpt.cpp:41-71) with the two random draws passed in.  Forces are
supplied by callables, in the tests the valence / electrostatics / vdW oracles.
"""
import numpy as np

EKCAL = 418.4                  # tinker/source/units.f:90
GASCONST = 1.9872042586e-3     # units.f:84


def kinetic(vel, mass, nfree):
    """eksum (kcal/mol) and temperature (K)."""
    eksum = 0.5 * float((mass[:, None] * vel * vel).sum()) / EKCAL
    return eksum, 2.0 * eksum / (nfree * GASCONST)


def bussi_scale(temp, dt, tautemp, kelvin, nfree, r, s):
    """Velocity scale of src/mdpt.cpp:54-61 for a normal deviate r and a chi-squared(nfree-1) deviate s."""
    if temp == 0:
        temp = 0.1
    c = np.exp(-dt / tautemp)
    d = (1.0 - c) * (kelvin / temp) / nfree
    scale = np.sqrt(c + (s + r * r) * d + 2.0 * r * np.sqrt(c * d))
    return -scale if r + np.sqrt(c / d) < 0 else scale


class Integrator:
    def __init__(self, xyz, vel, mass, fast_grad, slow_grad, dt, nrespa=1):
        """fast_grad(xyz), slow_grad(xyz) -> (n,3) gradients in kcal/mol/A; dt in ps.  Kick-off evaluates both."""
        self.x = np.array(xyz, float)
        self.v = np.array(vel, float)
        self.m = np.array(mass, float)
        self.minv = np.where(self.m > 0, 1.0 / np.where(self.m > 0, self.m, 1.0), 0.0)
        self.fast, self.slow = fast_grad, slow_grad
        self.dt, self.nrespa = float(dt), int(nrespa)
        self.gf = self.fast(self.x)
        self.gs = self.slow(self.x)

    def _kick(self, cf, cs):
        self.v += (-EKCAL * self.minv)[:, None] * (self.gf * cf + self.gs * cs)

    def step(self):
        dt, nr = self.dt, self.nrespa
        dta = dt / nr
        self._kick(0.5 * dta, 0.5 * dt)           # velR1(dt/2): fast over dt_a/2, slow over dt/2
        for _ in range(1, nr):
            self.x += dta * self.v                # pos(dta)
            self.gf = self.fast(self.x)
            self._kick(dta, 0.0)                  # velR0(dta)
        self.x += dta * self.v
        self.gf = self.fast(self.x)
        self.gs = self.slow(self.x)
        self._kick(0.5 * dta, 0.5 * dt)           # velR2
