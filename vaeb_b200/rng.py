"""Host-side noise sources.

`RandomStreams` re-creates the draw order of Theano's ``T.shared_randomstreams.RandomStreams``
as the reference uses it (VAEB.py:158,42): a seed generator ``RandomState(seed)``; every
``srng.normal`` node owns ``RandomState(gen.randint(2**30))``; each call of the compiled
function draws ``normal(0,1,shape)`` in fp64 and casts to floatX; ``update`` and ``validate``
share the node states.  This is recalled third-party behaviour (Theano is not vendored in the
reference and cannot run here) -- best effort, used only when ``eps_mode='theano'``.  The
default noise source is the on-device Philox generator (csrc/philox.cuh)."""
from __future__ import annotations

import numpy as np


class RandomStreams(object):
    def __init__(self, seed=10):
        self.gen_seedgen = np.random.RandomState(seed)
        self.nodes = []

    def new_node(self):
        self.nodes.append(np.random.RandomState(int(self.gen_seedgen.randint(2 ** 30))))
        return len(self.nodes) - 1

    def normal(self, node, shape):
        return self.nodes[node].normal(0.0, 1.0, size=shape).astype(np.float32)
