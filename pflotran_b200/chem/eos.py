"""Reference water density used for transport-only runs.

The reference sets `option%reference_water_density` from the IFC-67 equation
of state at the reference temperature/pressure (src/pflotran/
factory_subsurface.F90:995-1000) and copies it into every cell's `den_kg`
(init_subsurface_transport.F90:50).  `den_kg` enters molality<->molarity
conversions, so gold-file parity needs the same double.  Restated from
src/pflotran/eos_water.F90:876-1033 (density part only), same expression order.
"""
import math

H2O_CRITICAL_TEMPERATURE = 647.3      # pflotran_constants.F90:40
H2O_CRITICAL_PRESSURE = 22.064e6      # pflotran_constants.F90:44
FMWH2O = 18.01534                     # pflotran_constants.F90:33

_aa = [
    6.824687741e03, -5.422063673e02, -2.096666205e04, 3.941286787e04,
    -6.733277739e04, 9.902381028e04, -1.093911774e05, 8.590841667e04,
    -4.511168742e04, 1.418138926e04, -2.017271113e03, 7.982692717e00,
    -2.616571843e-2, 1.522411790e-3, 2.284279054e-2, 2.421647003e02,
    1.269716088e-10, 2.074838328e-7, 2.174020350e-8, 1.105710498e-9,
    1.293441934e01, 1.308119072e-5, 6.047626338e-14]
_a1, _a2, _a3, _a4 = 8.438375405e-1, 5.362162162e-4, 1.720000000e00, 7.342278489e-2
_a5, _a6, _a7, _a8 = 4.975858870e-2, 6.537154300e-1, 1.150000000e-6, 1.510800000e-5
_a9, _a10, _a11, _a12 = 1.418800000e-1, 7.002753165e00, 2.995284926e-4, 2.040000000e-1


def water_density_ifc67(t: float, p: float) -> float:
    """kg/m^3 at t [C], p [Pa]."""
    aa = _aa
    tc1 = H2O_CRITICAL_TEMPERATURE
    pc1 = H2O_CRITICAL_PRESSURE
    vc1 = 0.00317
    utc1 = 1.0 / tc1
    upc1 = 1.0 / pc1
    theta = (t + 273.15) * utc1
    theta2x = theta * theta
    theta18 = math.pow(theta, 18.0)
    theta20 = theta18 * theta2x
    beta = p * upc1
    beta2x = beta * beta
    yy = 1.0 - _a1 * theta2x - _a2 * math.pow(theta, -6.0)
    xx = _a3 * yy * yy - 2.0 * (_a4 * theta - _a5 * beta)
    if xx > 0.0:
        xx = math.sqrt(xx)
    else:
        xx = 1.0e-6
    zz = yy + xx
    u0 = -5.0 / 17.0
    u1 = aa[11] * _a5 * math.pow(zz, u0)
    u2 = 1.0 / (_a8 + math.pow(theta, 11.0))
    u3 = aa[17] + (2.0 * aa[18] + 3.0 * aa[19] * beta) * beta
    u4 = 1.0 / (_a7 + theta18 * theta)
    u5 = math.pow(_a10 + beta, -4.0)
    u6 = _a11 - 3.0 * u5
    u7 = aa[20] * theta18 * (_a9 + theta2x)
    u8 = aa[15] * math.pow(_a6 - theta, 9.0)
    vr = (u1 + aa[12] + theta * (aa[13] + aa[14] * theta) + u8 * (_a6 - theta)
          + aa[16] * u4 - u2 * u3 - u6 * u7
          + (3.0 * aa[21] * (_a12 - theta) + 4.0 * aa[22] * beta / theta20) * beta2x)
    return 1.0 / (vr * vc1)
