"""ScalarVector: the (scalar features, vector features) pair the reference passes between layers.

Mirrors the interface of ``ScalarVector`` in the reference (src/models/components/__init__.py:17-94):
a tuple subclass, so ``h, chi = layer(...)`` unpacks and ``isinstance(x, tuple)`` holds.
"""
from __future__ import annotations

import torch


class ScalarVector(tuple):
    def __new__(cls, scalar, vector):
        return super().__new__(cls, (scalar, vector))

    def __getnewargs__(self):
        return (self[0], self[1])

    scalar = property(lambda self: self[0])
    vector = property(lambda self: self[1])

    @staticmethod
    def _pair(other):
        return (other[0], other[1]) if isinstance(other, tuple) else (other.scalar, other.vector)

    def __add__(self, other):  # element-wise (comp/__init__.py:36-44), not tuple concatenation
        s, v = self._pair(other)
        return ScalarVector(self[0] + s, self[1] + v)

    def __mul__(self, other):
        if isinstance(other, tuple):
            return ScalarVector(self[0] * other[0], self[1] * other[1])
        return ScalarVector(self[0] * other, self[1] * other)

    def concat(self, others, dim=-1):
        dim %= self[0].dim()
        members = (self,) + tuple(others)
        return torch.cat([m[0] for m in members], dim=dim), torch.cat([m[1] for m in members], dim=dim)

    def flatten(self):
        """[..., s + 3c]: scalars first, then vectors channel-major with xyz fastest (comp:61-63)."""
        v = self[1]
        return torch.cat((self[0], v.reshape(v.shape[:-2] + (3 * v.shape[-2],))), dim=-1)

    @staticmethod
    def recover(x, vector_dim):
        n = 3 * vector_dim
        return ScalarVector(x[..., : x.shape[-1] - n], x[..., x.shape[-1] - n:].reshape(x.shape[:-1] + (vector_dim, 3)))

    def vs(self):
        return self[0], self[1]

    def idx(self, idx):
        return ScalarVector(self[0][idx], self[1][idx])

    def repeat(self, n, c=1, y=1):
        return ScalarVector(self[0].repeat(n, c), self[1].repeat(n, y, c))

    def clone(self):
        return ScalarVector(self[0].clone(), self[1].clone())

    def mask(self, node_mask):
        return ScalarVector(self[0] * node_mask[:, None], self[1] * node_mask[:, None, None])

    def __setitem__(self, key, value):
        self[0][key] = value[0]
        self[1][key] = value[1]

    def __repr__(self):
        return f"ScalarVector({self[0]}, {self[1]})"
