"""Unit conversion as the reference performs it (same floating-point steps).

Follows reference src/pflotran/units.F90:17-80 (UnitsConvertToInternal),
:84-180 (UnitsConvertParse: numerator/denominator split on '/'),
:183-312 (UnitsConvert: '-' separated products, user_to_SI / internal_to_SI)
and :384-514 (UnitsConvertToSI factors).  The order of multiplications and
divisions is kept so that e.g. 'cm^2/cm^3' -> 'm^2/m^3' gives the identical
double (1.d-4/1.d-6) the reference multiplies surface areas by.
"""

_SI = {
    'cm^3': 1.0e-6, 'ml': 1.0e-6, 'mL': 1.0e-6,
    'l': 1.0e-3, 'L': 1.0e-3, 'dm^3': 1.0e-3,
    'm^3': 1.0, 'gal': 3.785411784e-3, 'gallon': 3.785411784e-3,
    'cm^2': 1.0e-4, 'dm^2': 1.0e-2, 'm^2': 1.0, 'km^2': 1.0e6,
    'km': 1000.0, 'm': 1.0, 'met': 1.0, 'meter': 1.0, 'dm': 1.0e-1, 'cm': 1.0e-2, 'mm': 1.0e-3,
    's': 1.0, 'sec': 1.0, 'second': 1.0, 'min': 60.0, 'minute': 60.0,
    'h': 3600.0, 'hr': 3600.0, 'hour': 3600.0,
    'd': 24.0 * 3600.0, 'day': 24.0 * 3600.0,
    'w': 7.0 * 24.0 * 3600.0, 'week': 7.0 * 24.0 * 3600.0,
    'mo': 365.0 / 12.0 * 24.0 * 3600.0, 'month': 365.0 / 12.0 * 24.0 * 3600.0,
    'y': 365.0 * 24.0 * 3600.0, 'yr': 365.0 * 24.0 * 3600.0, 'year': 365.0 * 24.0 * 3600.0,
    'J': 1.0, 'kJ': 1.0e3, 'MJ': 1.0e6, 'W': 1.0, 'kW': 1.0e3, 'MW': 1.0e6,
    'mol': 1.0, 'mole': 1.0, 'moles': 1.0, 'kmol': 1.0e3,
    'ug': 1.0e-9, 'mg': 1.0e-6, 'g': 1.0e-3, 'kg': 1.0,
    'C': 1.0, 'Celsius': 1.0, 'Pa': 1.0, 'kPa': 1.0e3, 'MPa': 1.0e6, 'Bar': 1.0e5,
    'M': 1.0, 'mM': 1.0e-3, 'N': 1.0, 'unitless': 1.0, '1': 1.0,
}


def _convert(user: str, internal: str) -> float:
    conv_user = 1.0
    conv_int = 1.0
    for u in internal.split('-'):
        conv_int = conv_int * _SI[u]
    for u in user.split('-'):
        conv_user = conv_user * _SI[u]
    return conv_user / conv_int


def units_convert_to_internal(units: str, internal_units: str) -> float:
    if ('/' in units) != ('/' in internal_units):
        raise ValueError('unit structure mismatch: %s vs %s' % (units, internal_units))
    if '/' in units:
        un, ud = units.split('/', 1)
        inn, ind = internal_units.split('/', 1)
        return _convert(un, inn) / _convert(ud, ind)
    return _convert(units, internal_units)
