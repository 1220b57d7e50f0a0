"""Host-side mirror of the reference's univariate polynomials (code/univariate.py).

Only used when the reference's ``univariate`` module is not importable (see
algebra.py in this directory).  Coefficient-list semantics are the reference's:
the constructor copies the list (:4-5), ``degree`` ignores trailing zeros and is
-1 for the zero polynomial (:7-17), ``*`` is schoolbook and returns
len(a)+len(b)-1 coefficients untrimmed (:38-48), ``/`` asserts a zero remainder
(:50-53), ``scale`` multiplies coefficient i by factor^i (:153-154).  The heavy
loops work on Python ints and wrap the results, which keeps values identical
while being much faster than element-object arithmetic.
"""
from algebra import *  # noqa: F401,F403  (re-exported, like the reference does)
from algebra import FieldElement


def _field_of(coefficients):
    return coefficients[0].field


class Polynomial:
    def __init__(self, coefficients):
        self.coefficients = [c for c in coefficients]

    # ------------------------------------------------------------ inspection
    def degree(self):
        d = -1
        for i, c in enumerate(self.coefficients):
            if c.value != 0:
                d = i
        return d

    def is_zero(self):
        return all(c.value == 0 for c in self.coefficients)

    def leading_coefficient(self):
        return self.coefficients[self.degree()]

    def __eq__(self, other):
        if self.degree() != other.degree():
            return False
        if self.degree() == -1:
            return True
        return all(self.coefficients[i] == other.coefficients[i] for i in range(len(self.coefficients)))

    def __neq__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __str__(self):
        return "[" + ",".join(str(c) for c in self.coefficients) + "]"

    # ------------------------------------------------------------ arithmetic
    def __neg__(self):
        return Polynomial([-c for c in self.coefficients])

    def __add__(self, other):
        if self.degree() == -1:
            return other
        if other.degree() == -1:
            return self
        field = _field_of(self.coefficients)
        p = field.p
        n = max(len(self.coefficients), len(other.coefficients))
        acc = [0] * n
        for i, c in enumerate(self.coefficients):
            acc[i] = c.value
        for i, c in enumerate(other.coefficients):
            acc[i] = (acc[i] + c.value) % p
        return Polynomial([FieldElement(v, field) for v in acc])

    def __sub__(self, other):
        return self.__add__(-other)

    def __mul__(self, other):
        if self.coefficients == [] or other.coefficients == []:
            return Polynomial([])
        field = _field_of(self.coefficients)
        p = field.p
        a = [c.value for c in self.coefficients]
        b = [c.value for c in other.coefficients]
        acc = [0] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            if x == 0:
                continue
            for j, y in enumerate(b):
                acc[i + j] += x * y
        return Polynomial([FieldElement(v % p, field) for v in acc])

    def divide(numerator, denominator):
        """Long division -> (quotient, remainder); None for a zero denominator (:80-97)."""
        dd = denominator.degree()
        if dd == -1:
            return None
        dn = numerator.degree()
        if dn < dd:
            return (Polynomial([]), numerator)
        field = _field_of(denominator.coefficients)
        p = field.p
        rem = [c.value for c in numerator.coefficients]
        den = [c.value for c in denominator.coefficients[:dd + 1]]
        lead_inv = pow(den[dd], -1, p)
        quo = [0] * (dn - dd + 1)
        for top in range(dn, dd - 1, -1):
            if rem[top] == 0:
                continue
            coefficient = rem[top] * lead_inv % p
            shift = top - dd
            quo[shift] = coefficient
            for j, y in enumerate(den):
                rem[shift + j] = (rem[shift + j] - coefficient * y) % p
        return (Polynomial([FieldElement(v, field) for v in quo]),
                Polynomial([FieldElement(v, field) for v in rem]))

    def __truediv__(self, other):
        quo, rem = Polynomial.divide(self, other)
        assert rem.is_zero(), "cannot perform polynomial division because remainder is not zero"
        return quo

    def __mod__(self, other):
        quo, rem = Polynomial.divide(self, other)
        return rem

    def __xor__(self, exponent):
        if self.is_zero():
            return Polynomial([])
        one = Polynomial([_field_of(self.coefficients).one()])
        if exponent == 0:
            return one
        acc = one
        for bit in bin(exponent)[2:]:
            acc = acc * acc
            if bit == "1":
                acc = acc * self
        return acc

    # ------------------------------------------------------------ evaluation
    def evaluate(self, point):
        p = point.field.p
        x = point.value
        acc = 0
        for c in reversed(self.coefficients):
            acc = (acc * x + c.value) % p
        return FieldElement(acc, point.field)

    def evaluate_domain(self, domain):
        return [self.evaluate(d) for d in domain]

    def scale(self, factor):
        p = factor.field.p
        out, power = [], 1
        for c in self.coefficients:
            out.append(FieldElement(c.value * power % p, c.field))
            power = power * factor.value % p
        return Polynomial(out)

    # ------------------------------------------------------ lagrange helpers
    def zerofier_domain(domain):
        """prod (X - d) over the domain (:122-128), built on ints."""
        field = domain[0].field
        p = field.p
        acc = [1]
        for d in domain:
            nxt = [0] * (len(acc) + 1)
            for i, c in enumerate(acc):
                nxt[i] = (nxt[i] - c * d.value) % p
                nxt[i + 1] = (nxt[i + 1] + c) % p
            acc = nxt
        return Polynomial([FieldElement(v, field) for v in acc])

    def interpolate_domain(domain, values):
        """Lagrange interpolation (:107-120).  O(n^2): divide the zerofier by each
        (X - d_i) synthetically instead of multiplying n-1 linear factors per point;
        the interpolant (n coefficients) is the same polynomial."""
        assert len(domain) == len(values), \
            "number of elements in domain does not match number of values -- cannot interpolate"
        assert len(domain) > 0, "cannot interpolate between zero points"
        field = domain[0].field
        p = field.p
        n = len(domain)
        z = [c.value for c in Polynomial.zerofier_domain(domain).coefficients]
        acc = [0] * n
        for i in range(n):
            d = domain[i].value
            # q = z / (X - d) by synthetic division, highest coefficient first
            q = [0] * n
            carry = 0
            for k in range(n, 0, -1):
                carry = (z[k] + carry * d) % p
                q[k - 1] = carry
            denom = 0
            for c in reversed(q):
                denom = (denom * d + c) % p
            weight = values[i].value * pow(denom, -1, p) % p
            if weight:
                for k in range(n):
                    acc[k] = (acc[k] + weight * q[k]) % p
        return Polynomial([FieldElement(v, field) for v in acc])


def test_colinearity(points):
    xs = [pt[0] for pt in points]
    ys = [pt[1] for pt in points]
    return Polynomial.interpolate_domain(xs, ys).degree() == 1
