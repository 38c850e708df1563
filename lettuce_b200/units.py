"""Physical <-> lattice unit conversion with the API of lettuce's `UnitConversion` (lettuce/_unit.py:13-145).

Every conversion is a rescaling by the ratio of a characteristic quantity in the two unit systems, so the
class is generated from a table: quantity -> (characteristic scale, rounding order).  The rounding order is
part of the contract -- `convert_length_to_pu` feeds the obstacle masks (obstacle.py:101-105) and the
initial conditions, and parity with the reference's fields needs the same sequence of multiplications and
divisions: most quantities are converted as `x / from * to`, lengths and energies as `x * to / from`.
Conversions accept floats, numpy arrays and torch tensors.
"""
from math import sqrt

__all__ = ["UnitConversion"]

_DIVIDE_FIRST, _MULTIPLY_FIRST = 0, 1

# quantity -> (name of the characteristic-scale accessor without the _lu/_pu suffix, operation order)
_QUANTITIES = {
    "velocity": ("characteristic_velocity", _DIVIDE_FIRST),
    "acceleration": ("_characteristic_acceleration", _DIVIDE_FIRST),
    "time": ("_characteristic_time", _DIVIDE_FIRST),
    "density": ("characteristic_density", _DIVIDE_FIRST),
    "pressure": ("characteristic_pressure", _DIVIDE_FIRST),
    "length": ("characteristic_length", _MULTIPLY_FIRST),
    "energy": ("characteristic_pressure", _MULTIPLY_FIRST),              # energy density = density * velocity^2
    "incompressible_energy": ("_characteristic_velocity_squared", _MULTIPLY_FIRST),
}


class UnitConversion:
    def __init__(self, reynolds_number, mach_number=0.05, characteristic_length_pu=1, characteristic_velocity_pu=1,
                 characteristic_length_lu=1, characteristic_density_lu=1, characteristic_density_pu=1,
                 cs=1 / sqrt(3.0)):
        self.cs = cs
        self.reynolds_number = reynolds_number
        self.mach_number = mach_number
        self.characteristic_length_pu = characteristic_length_pu
        self.characteristic_velocity_pu = characteristic_velocity_pu
        self.characteristic_length_lu = characteristic_length_lu
        self.characteristic_density_lu = characteristic_density_lu
        self.characteristic_density_pu = characteristic_density_pu

    # ---- characteristic scales (the lattice velocity follows from the Mach number) ----------------------
    @property
    def characteristic_velocity_lu(self):
        return self.cs * self.mach_number

    def _scale(self, name, system):
        return getattr(self, f"{name}_{system}")

    def __getattr__(self, name):
        # derived scales in both unit systems, computed from the primary ones on demand
        for system in ("lu", "pu"):
            if not name.endswith("_" + system):
                continue
            base = name[:-3]
            length = self.__dict__.get("characteristic_length_" + system)
            density = self.__dict__.get("characteristic_density_" + system)
            if length is None or density is None:
                break
            velocity = self.characteristic_velocity_lu if system == "lu" else self.__dict__["characteristic_velocity_pu"]
            if base == "characteristic_pressure":
                return density * velocity ** 2
            if base == "_characteristic_time":
                return length / velocity
            if base == "_characteristic_acceleration":
                return velocity ** 2 / length
            if base == "_characteristic_velocity_squared":
                return velocity ** 2
            if base == "viscosity":
                return length * velocity / self.reynolds_number
        raise AttributeError(name)

    @property
    def relaxation_parameter_lu(self):
        """tau = nu / cs^2 + 1/2 (lettuce/_unit.py:58-60)"""
        return self.viscosity_lu / self.cs ** 2 + 0.5

    # ---- pressure <-> density of the weakly compressible model (lettuce/_unit.py:94-101) ----------------
    def convert_density_lu_to_pressure_pu(self, density_lu):
        return self.convert_pressure_to_pu((density_lu - self.characteristic_density_lu) * self.cs ** 2)

    def convert_pressure_pu_to_density_lu(self, pressure_pu):
        return self.convert_pressure_to_lu(pressure_pu) / self.cs ** 2 + self.characteristic_density_lu


def _make_converter(quantity, scale, order, source, target):
    def convert(self, value):
        a, b = self._scale(scale, source), self._scale(scale, target)
        return value / a * b if order == _DIVIDE_FIRST else value * b / a
    convert.__name__ = f"convert_{quantity}_to_{target}"
    convert.__doc__ = f"{quantity} given in {source} expressed in {target}"
    return convert


for _quantity, (_scale_name, _order) in _QUANTITIES.items():
    setattr(UnitConversion, f"convert_{_quantity}_to_pu", _make_converter(_quantity, _scale_name, _order, "lu", "pu"))
    setattr(UnitConversion, f"convert_{_quantity}_to_lu", _make_converter(_quantity, _scale_name, _order, "pu", "lu"))
