"""Host-side mirror of the reference's Merkle tree (code/merkle.py:3-44).

Only used when the reference's ``merkle`` module is not importable.  Iterative
level-by-level tree instead of the reference's recursion; same hashes:
leaf = blake2b(bytes(element)), node = blake2b(left + right), paths are the
siblings bottom-up.  (The drop-in ``fri`` module replaces commit/open with the
GPU tree; this class is the host verifier and the CPU cross-check in tests.)
"""
from hashlib import blake2b


class Merkle:
    H = blake2b

    def _levels(leafs):
        assert len(leafs) & (len(leafs) - 1) == 0, "length must be power of two"
        levels = [list(leafs)]
        while len(levels[-1]) > 1:
            cur = levels[-1]
            levels.append([Merkle.H(cur[i] + cur[i + 1]).digest() for i in range(0, len(cur), 2)])
        return levels

    def commit_(leafs):
        return Merkle._levels(leafs)[-1][0]

    def commit(data_array):
        return Merkle.commit_([Merkle.H(bytes(da)).digest() for da in data_array])

    def open_(index, leafs):
        assert len(leafs) & (len(leafs) - 1) == 0, "length must be power of two"
        assert 0 <= index and index < len(leafs), "cannot open invalid index"
        path = []
        for level in Merkle._levels(leafs)[:-1]:
            path.append(level[index ^ 1])
            index >>= 1
        return path

    def open(index, data_array):
        return Merkle.open_(index, [Merkle.H(bytes(da)).digest() for da in data_array])

    def verify_(root, index, path, leaf):
        assert 0 <= index and index < (1 << len(path)), "cannot verify invalid index"
        node = leaf
        for sibling in path:
            node = Merkle.H(node + sibling).digest() if index % 2 == 0 else Merkle.H(sibling + node).digest()
            index >>= 1
        return root == node

    def verify(root, index, path, data_element):
        return Merkle.verify_(root, index, path, Merkle.H(bytes(data_element)).digest())
