"""Host-side mirror of the reference's proof stream (code/ip.py:4-30).

Only used when the reference's ``ip`` module is not importable (see algebra.py in
this directory).  Challenges are shake_256 over pickle.dumps of the object list,
exactly as the reference computes them, so a transcript built here and one built
by the reference agree byte for byte.
"""
import pickle as pickle
from hashlib import shake_256


def _challenge(objects, num_bytes):
    """Fiat-Shamir: squeeze `num_bytes` out of shake_256(pickle(objects))  (ip.py:21-25)"""
    return shake_256(pickle.dumps(objects)).digest(num_bytes)


class ProofStream:
    """A list-backed transcript: the prover appends, the verifier replays with a cursor."""

    def __init__(self):
        self.objects = []      # everything pushed so far, in order
        self.read_index = 0    # verifier cursor

    # -- prover side ------------------------------------------------------
    def push(self, obj):
        self.objects.append(obj)

    def serialize(self):
        return pickle.dumps(self.objects)

    def prover_fiat_shamir(self, num_bytes=32):
        # the prover hashes the whole transcript
        return _challenge(self.objects, num_bytes)

    # -- verifier side ----------------------------------------------------
    def pull(self):
        assert self.read_index < len(self.objects), "ProofStream: cannot pull object; queue empty."
        self.read_index += 1
        return self.objects[self.read_index - 1]

    def verifier_fiat_shamir(self, num_bytes=32):
        # the verifier hashes only what it has read so far
        return _challenge(self.objects[:self.read_index], num_bytes)

    def deserialize(self, bb):
        fresh = ProofStream()
        fresh.objects = pickle.loads(bb)
        return fresh
