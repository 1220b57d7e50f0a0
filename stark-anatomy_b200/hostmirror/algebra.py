"""Host-side mirror of the reference's field value types (code/algebra.py).

Used ONLY when the reference's own ``algebra`` module is not importable (the
GPU test box has no /root/reference): the drop-in ``ntt`` / ``fri`` modules
import ``algebra`` by name, so a user running the real reference keeps the
reference's classes, and this file steps in otherwise.

The module and class names are the reference's on purpose: the Fiat-Shamir
transcript is ``shake_256(pickle.dumps(objects))`` (code/ip.py:18-22) and pickle
records ``algebra.FieldElement`` / ``algebra.Field`` plus the instance dict in
insertion order (value, field).  tests/test_hostmirror.py checks that pickles
made with these classes are byte-identical to the reference's.

Semantics mirrored (file:line in /root/reference/code/algebra.py):
  canonical residues in [0, p); add :78-79, subtract :81-82, multiply :75-76,
  negate :84-85, inverse via extended Euclid with inverse(0) == 0 :87-89,
  divide asserts "divide by zero" :91-94, ``^`` is exponentiation :38-45,
  ``==`` compares values only :47-48, bytes() is decimal ASCII :53-57,
  Field.main / generator / primitive_nth_root / sample :96-120.
"""

_MAIN_P = 1 + 407 * (1 << 119)
_MAIN_GENERATOR = 85408008396924667383611388730472331217


def xgcd(x, y):
    """Extended Euclid: returns (a, b, g) with a*x + b*y == g (algebra.py:1-12)."""
    r0, r1 = x, y
    s0, s1 = 1, 0
    t0, t1 = 0, 1
    while r1:
        q, rem = divmod(r0, r1)
        r0, r1 = r1, rem
        s0, s1 = s1, s0 - q * s1
        t0, t1 = t1, t0 - q * t1
    return s0, t0, r0


class FieldElement:
    def __init__(self, value, field):
        self.value = value
        self.field = field

    # arithmetic is delegated to the field, like the reference does
    def __add__(self, other):
        return self.field.add(self, other)

    def __sub__(self, other):
        return self.field.subtract(self, other)

    def __mul__(self, other):
        return self.field.multiply(self, other)

    def __truediv__(self, other):
        return self.field.divide(self, other)

    def __neg__(self):
        return self.field.negate(self)

    def inverse(self):
        return self.field.inverse(self)

    def __xor__(self, exponent):
        """Modular exponentiation (algebra.py:38-45); same value as square-and-multiply."""
        return FieldElement(pow(self.value, exponent, self.field.p) if exponent else 1 % self.field.p,
                            self.field)

    def __eq__(self, other):
        return self.value == other.value

    def __neq__(self, other):
        return self.value != other.value

    __hash__ = None

    def __str__(self):
        return str(self.value)

    def __bytes__(self):
        return str(self.value).encode()

    def __repr__(self):
        return "FieldElement(%d)" % self.value

    def is_zero(self):
        return self.value == 0


class Field:
    def __init__(self, p):
        self.p = p

    def zero(self):
        return FieldElement(0, self)

    def one(self):
        return FieldElement(1, self)

    def add(self, left, right):
        return FieldElement((left.value + right.value) % self.p, self)

    def subtract(self, left, right):
        return FieldElement((left.value - right.value) % self.p, self)

    def multiply(self, left, right):
        return FieldElement(left.value * right.value % self.p, self)

    def negate(self, operand):
        return FieldElement(-operand.value % self.p, self)

    def inverse(self, operand):
        a, _, _ = xgcd(operand.value, self.p)
        return FieldElement(a % self.p, self)

    def divide(self, left, right):
        assert not right.is_zero(), "divide by zero"
        a, _, _ = xgcd(right.value, self.p)
        return FieldElement(left.value * a % self.p, self)

    def main():
        return Field(_MAIN_P)

    def generator(self):
        assert self.p == _MAIN_P, "Do not know generator for other fields beyond 1+407*2^119"
        return FieldElement(_MAIN_GENERATOR, self)

    def primitive_nth_root(self, n):
        assert self.p == _MAIN_P, "Unknown field, can't return root of unity."
        assert n <= 1 << 119 and (n & (n - 1)) == 0, \
            "Field does not have nth root of unity where n > 2^119 or not power of two."
        # the generator has order 2^119: raise it to 2^119 / n
        return FieldElement(pow(_MAIN_GENERATOR, (1 << 119) // n, self.p), self)

    def sample(self, byte_array):
        return FieldElement(int.from_bytes(bytes(byte_array), "big") % self.p, self)
