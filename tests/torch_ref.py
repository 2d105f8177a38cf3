"""Independent torch-float64 autograd restatement of the hot-path ops, used ONLY to validate the NumPy oracle
("independent torch-CPU float64 autograd re-implementation must agree").  It deliberately shares
no code with oracle/ops.py: the LSTM is written step-by-step with a custom grad-clip autograd Function, the
delta operator as explicit gathers, the losses straight from the reference formulas."""
import torch


class GradClip(torch.autograd.Function):
    """theano.gradient.grad_clip: identity forward, clamp of the incoming gradient backward."""

    @staticmethod
    def forward(ctx, x, lo, hi):
        ctx.lo, ctx.hi = lo, hi
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.clamp(ctx.lo, ctx.hi), None, None


def delta_coeff(x, theta):
    N, T, F = x.shape
    t = torch.arange(T)
    d = torch.zeros_like(x)
    for th in range(1, theta + 1):
        hi = x[:, torch.clamp(t + th, max=T - 1)]
        lo = x[:, torch.clamp(t - th, min=0)]
        d = d + th * (hi - lo) / (2.0 * th * th)
    return d


def delta_layer(x, theta):
    d = delta_coeff(x, theta)
    a = delta_coeff(d, theta)
    return torch.cat([x, d, a], dim=2)


def lstm(x, mask, W_in, W_hid, b, peep, cell_init, hid_init, backwards, clip=5.0):
    N, T, I = x.shape
    H = W_hid.shape[0]
    xW = x.reshape(N * T, I) @ W_in + b
    xW = xW.reshape(N, T, 4 * H)
    c = cell_init.reshape(1, H).expand(N, H)
    h = hid_init.reshape(1, H).expand(N, H)
    outs = [None] * T
    order = range(T - 1, -1, -1) if backwards else range(T)
    for t in order:
        g = xW[:, t] + h @ W_hid
        if clip:
            g = GradClip.apply(g, -clip, clip)
        gi, gf, gc, go = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        if peep is not None:
            gi = gi + c * peep[0]
            gf = gf + c * peep[1]
        i, f, cin = torch.sigmoid(gi), torch.sigmoid(gf), torch.tanh(gc)
        cn = f * c + i * cin
        if peep is not None:
            go = go + cn * peep[2]
        o = torch.sigmoid(go)
        hn = o * torch.tanh(cn)
        m = mask[:, t].bool().unsqueeze(1)
        c = torch.where(m, cn, c)
        h = torch.where(m, hn, h)
        outs[t] = h
    return torch.stack(outs, dim=1)


def temporal_softmax_loss(probs, y, mask):
    N, T, C = probs.shape
    x = probs.reshape(N * T, C)
    q = torch.softmax(x, dim=1)
    mf = mask.reshape(N * T).to(x.dtype)
    picked = q[torch.arange(N * T), y.reshape(N * T).long()]
    return -(mf * torch.log(picked)).sum() / mf.sum()
