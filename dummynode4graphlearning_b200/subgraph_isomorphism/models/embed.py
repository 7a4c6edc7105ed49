"""Encoders / embeddings on the input side of the counting models.

Mirrors ``subgraph_isomorphism/models/embed.py``: ``MultihotEmbedding`` (:197-208, frozen 0/1 table of
fixed-width base-b digits, MSB first, ``_int2multihot`` :70-87) and the ``Embedding`` family (:105-194) whose
``forward`` accepts either integer ids (table lookup) or an already-encoded float matrix (``enc @ weight``).
Tables are tiny (K <= 14 columns); the GEMM is plain library work, not a hot kernel.
"""
import torch as th
import torch.nn as nn


def get_enc_len(x, base=10):
    """number of base-`base` digits of x (>= 1)   (embed.py:8-35)"""
    x, cnt = int(x), 0
    while x > 0:
        x //= base
        cnt += 1
    return max(cnt, 1)


def int2multihot(values, len_x, base):
    """(len(values), len_x*base) 0/1 matrix: digit d at position i (MSB first) sets column i*base+d."""
    v = th.as_tensor(values, dtype=th.long) % (base ** len_x)
    rep = th.zeros((v.numel(), len_x * base), dtype=th.float32)
    for pos in range(len_x - 1, -1, -1):
        rep[th.arange(v.numel()), pos * base + (v % base)] = 1.0
        v = v // base
    return rep


class Embedding(nn.Embedding):
    def forward(self, x):
        if x.dtype == th.long:
            return super().forward(x)
        if x.dtype == th.float and x.size(-1) == self.num_embeddings:
            lead = x.size()[:-1]
            return th.matmul(x.reshape(-1, x.size(-1)), self.weight).view(lead + (self.embedding_dim,))
        raise NotImplementedError

    def get_output_dim(self):
        return self.embedding_dim


class NormalEmbedding(Embedding):
    def __init__(self, num_embeddings, embedding_dim, **kw):
        super().__init__(num_embeddings, embedding_dim, **kw)
        nn.init.normal_(self.weight, 0.0, 1.0)


class UniformEmbedding(Embedding):
    def __init__(self, num_embeddings, embedding_dim, **kw):
        super().__init__(num_embeddings, embedding_dim, **kw)
        nn.init.uniform_(self.weight, -1.0, 1.0)


class OrthogonalEmbedding(Embedding):
    def __init__(self, num_embeddings, embedding_dim, **kw):
        super().__init__(num_embeddings, embedding_dim, **kw)
        nn.init.orthogonal_(self.weight)


class EquivariantEmbedding(Embedding):
    """circulant initialisation from ``row_vec`` (embed.py:162-173).  ``row_vec`` is a Parameter that
    never receives a gradient in the reference (its ``backward`` hook is never invoked, SURVEY.md
    App. A-13); the same holds here (``p.grad is None``)."""

    def __init__(self, num_embeddings, embedding_dim, **kw):
        super().__init__(num_embeddings, embedding_dim, **kw)
        self.row_vec = nn.Parameter(th.Tensor(self.embedding_dim))
        nn.init.normal_(self.row_vec, 0.0, 1.0)
        with th.no_grad():
            for i in range(num_embeddings):
                self.weight[i].copy_(th.roll(self.row_vec, i, 0))


class MultihotEmbedding(Embedding):
    def __init__(self, max_n=1024, base=2):
        self.max_n, self.base = max_n, base
        enc_len = get_enc_len(max_n - 1, base)
        super().__init__(max_n, base * enc_len)   # the reference writes 2*enc_len (embed.py:203): base is always 2 there
        with th.no_grad():
            self.weight.copy_(int2multihot(th.arange(max_n), enc_len, base))

    def extra_repr(self):
        return "base=%d, max_n=%d, enc_dim=%d" % (self.base, self.max_n, self.weight.shape[1])


class PositionEmbedding(Embedding):
    """frozen sinusoid table of ``--enc_net Position`` (embed.py:211-222): row i = [sin(i * f_k) ..., cos(i * f_k) ...]
    with f_k = 10000^(-2k / embedding_dim).  Same width as the multi-hot table it replaces
    (``get_enc_len(max_n - 1, base) * base`` columns, basemodel.py:642-646), so every downstream shape is unchanged."""

    def __init__(self, embedding_dim, max_len=512, scale=1):
        freq_seq = th.arange(0, embedding_dim, 2.0, dtype=th.float)
        inv_freq = th.pow(10000, (freq_seq / embedding_dim)).reciprocal()
        sinusoid_inp = th.ger(th.arange(0, max_len, 1.0), inv_freq)
        super().__init__(max_len, embedding_dim)
        with th.no_grad():
            self.weight.copy_(th.cat([th.sin(sinusoid_inp), th.cos(sinusoid_inp)], dim=-1) * scale)

    def extra_repr(self):
        return "embedding_dim=%d, max_len=%d" % (self.weight.shape[1], self.weight.shape[0])
