"""Host-side mirror of the reference's proof stream (code/ip.py:4-30).

Only used when the reference's ``ip`` module is not importable (see algebra.py in
this directory).  Challenges are shake_256 over pickle.dumps of the object list,
exactly as the reference computes them, so a transcript built here and one built
by the reference agree byte for byte.
"""
from hashlib import shake_256
import pickle as pickle


class ProofStream:
    def __init__(self):
        self.objects = []
        self.read_index = 0

    def push(self, obj):
        self.objects.append(obj)

    def pull(self):
        assert self.read_index < len(self.objects), "ProofStream: cannot pull object; queue empty."
        obj = self.objects[self.read_index]
        self.read_index += 1
        return obj

    def serialize(self):
        return pickle.dumps(self.objects)

    def prover_fiat_shamir(self, num_bytes=32):
        return shake_256(self.serialize()).digest(num_bytes)

    def verifier_fiat_shamir(self, num_bytes=32):
        seen = self.objects[:self.read_index]
        return shake_256(pickle.dumps(seen)).digest(num_bytes)

    def deserialize(self, bb):
        ps = ProofStream()
        ps.objects = pickle.loads(bb)
        return ps
