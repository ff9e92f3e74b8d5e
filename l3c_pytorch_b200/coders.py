"""ArithmeticCoder: the thin per-stream wrapper of the reference
(/root/reference/src/bitcoding/coders.py:33-90) over the torchac drop-in.  The batched path in
codec.py does not go through this class; it exists so that code written against the reference's
`range_encode` / `range_decode` keeps working."""
import torch

from . import torchac
from .dmll import CDFOut
from .times import NoOp


class ArithmeticCoder(object):
    def __init__(self, L):
        self.L = L

    def range_encode(self, data, cdf, time_logger=NoOp):
        """data: int16 [1,H,W] (any device); cdf: CDFOut or int16 NHWLp table -> bytes."""
        assert len(data.shape) == 3, data.shape
        with time_logger.run('data -> cpu'):
            data = data.to('cpu')
        assert data.dtype == torch.int16, 'Wrong dtype: {}'.format(data.dtype)
        data = data.reshape(-1).contiguous()
        if isinstance(cdf, CDFOut):
            pi, mu, ls, _, targets = cdf
            with time_logger.run('ac.encode'):
                return torchac.encode_logistic_mixture(targets, mu, ls, pi, data)
        N, H, W, Lp = cdf.shape
        assert Lp == self.L + 1, (Lp, self.L)
        with time_logger.run('ac.encode'):
            return torchac.encode_cdf(cdf, data)

    def range_decode(self, encoded_bytes, cdf, time_logger=NoOp):
        """-> int16 [N,H,W] on CPU."""
        if isinstance(cdf, CDFOut):
            pi, mu, ls, _, targets = cdf
            N, _, H, W = mu.shape
            with time_logger.run('ac.encode'):
                decoded = torchac.decode_logistic_mixture(targets, mu, ls, pi, encoded_bytes)
        else:
            N, H, W, Lp = cdf.shape
            assert Lp == self.L + 1, (Lp, self.L)
            with time_logger.run('ac.encode'):
                decoded = torchac.decode_cdf(cdf, encoded_bytes)
        return decoded.reshape(N, H, W)


class CodingCDFNonshared(object):
    """coders_helpers.py:31-56: hands out the CDF parameters channel by channel."""

    def __init__(self, l, total_C, dmll):
        self.l, self.dmll, self.total_C = l, dmll, total_C
        self.targets = dmll.targets(l.device)
        self.c_cur = 0

    def get_next_C(self, decoded_x):
        out = self.dmll.cdf_step_non_shared(self.l, self.targets, self.c_cur, self.total_C, decoded_x)
        self.c_cur += 1
        return out
