"""Log-cell parsing of the host mirror (AutoStopMCMC.java:460-466: Double.parseDouble, NumberFormatException -> 0.0)."""
import locale
import math

import pytest

from asm_b200 import criteria as K

# token -> what Double.parseDouble returns (None: NumberFormatException, the runner stores 0.0)
JAVA = {
    "1.5": 1.5, "-1.5": -1.5, "+1.5": 1.5, "1.": 1.0, ".5": 0.5, "-.5e1": -5.0, "1e3": 1000.0, "1E-3": 0.001, "1e+3": 1000.0,
    "007": 7.0, "1.5d": 1.5, "1.5D": 1.5, "2f": 2.0, "2F": 2.0, "1e3f": 1000.0, " 3.25 ": 3.25, "\t4\n": 4.0,
    "NaN": math.nan, "-NaN": math.nan, "+NaN": math.nan, "Infinity": math.inf, "+Infinity": math.inf, "-Infinity": -math.inf,
    "0x1p3": 8.0, "0X1.8P1": 3.0, "-0x.8p1": -1.0, "0x1p3d": 8.0,
    "1e999": math.inf, "-1e999": -math.inf, "1e-999": 0.0, "4.9e-324": 4.9e-324,
    "-7342.518290283": -7342.518290283, "0.1": 0.1, "123456789012345678901234567890": 1.2345678901234568e29,
    # NumberFormatException in Java; C's strtod would have accepted several of these
    "inf": None, "-inf": None, "+inf": None, "nan": None, "+nan": None, "-nan": None, "infinity": None, "-infinity": None,
    "INF": None, "NAN": None, "NaNf": None, "Infinityd": None, "0x10": None, "0x1.8": None, "0xp3": None, "1e": None, "1e+": None,
    ".": None, "-": None, "+": None, "": None, "e5": None, "1,5": None, "1.5x": None, "1 2": None, "--1": None, "1.2.3": None,
    "1d5": None, "abc": None, "1_000": None, "0b101": None, "#": None,
}


@pytest.mark.parametrize("token", sorted(JAVA))
def test_log_cell_follows_double_parse_double(token):
    want = JAVA[token]
    v, parsed = K.parse_log_cell(token)
    if want is None:
        assert not parsed and v == 0.0 and math.copysign(1.0, v) == 1.0
    elif math.isnan(want):
        assert parsed and math.isnan(v)
    else:
        assert parsed and v == want and math.copysign(1.0, v) == math.copysign(1.0, want)


def test_log_cell_ignores_the_process_locale():
    # an embedding process with a comma-decimal LC_NUMERIC must not change how "1.5" reads
    old = locale.setlocale(locale.LC_NUMERIC)
    for name in ("de_DE.UTF-8", "fr_FR.UTF-8", "de_DE", "nl_NL.UTF-8"):
        try:
            locale.setlocale(locale.LC_NUMERIC, name)
            break
        except locale.Error:
            continue
    else:
        pytest.skip("no comma-decimal locale installed")
    try:
        assert K.parse_log_cell("1.5") == (1.5, True)
        assert K.parse_log_cell("1,5") == (0.0, False)
    finally:
        locale.setlocale(locale.LC_NUMERIC, old)
