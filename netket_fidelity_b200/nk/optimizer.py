"""Sgd / Adam (NetKet `nk.optimizer.Sgd`, `nk.optimizer.Adam` = optax).  The update itself is a
CUDA kernel (nkfb_optimizer_update) acting on the flat parameter buffer of an MCState."""
import torch


class _Opt:
    kind = 0

    def __init__(self, learning_rate):
        self.learning_rate = float(learning_rate)
        self._state = {}

    def apply(self, vstate, flat_grad):
        """In-place update of vstate's flat parameter buffer with the flat gradient (P,) complex."""
        raise NotImplementedError


class Sgd(_Opt):
    kind = 0

    def apply(self, vstate, flat_grad):
        vstate._h.optimizer_update(0, vstate._flat, flat_grad, None, None, self.learning_rate, 0.0, 0.0, 0.0, 1,
                                   not vstate.model.is_complex)
        vstate._params_changed()

    def __repr__(self):
        return f"Sgd(learning_rate={self.learning_rate})"


class Adam(_Opt):
    kind = 1

    def __init__(self, learning_rate=0.001, b1=0.9, b2=0.999, eps=1e-8):
        super().__init__(learning_rate)
        self.b1, self.b2, self.eps = b1, b2, eps

    def apply(self, vstate, flat_grad):
        st = self._state.get(id(vstate))
        if st is None:
            st = dict(m=torch.zeros_like(vstate._flat),
                      v=torch.zeros(vstate._flat.shape, dtype=vstate._flat.real.dtype, device=vstate._flat.device),
                      t=0)
            self._state[id(vstate)] = st
        st["t"] += 1
        vstate._h.optimizer_update(1, vstate._flat, flat_grad, st["m"], st["v"], self.learning_rate, self.b1,
                                   self.b2, self.eps, st["t"], not vstate.model.is_complex)
        vstate._params_changed()

    def __repr__(self):
        return f"Adam(learning_rate={self.learning_rate}, b1={self.b1}, b2={self.b2}, eps={self.eps})"


class SR:
    """Stochastic-reconfiguration preconditioner, NetKet `nk.optimizer.SR(diag_shift=0.01)` with the on-the-fly
    quantum geometric tensor (`QGTOnTheFly`) and a conjugate-gradient solve -- the object the reference's driver
    takes as `preconditioner=` (driver/infidelity_optimizer.py:40,148,205-233):

        (S + diag_shift) dp = grad,    S_kl = <O_k^* O_l> - <O_k^*><O_l>   over the samples of the state,

    O_k = d log psi / d theta_k.  S is never formed (P = 40 500 at the north-star shape): one product S v is the
    Jacobian-vector product u_s = sum_l O_l(x_s) v_l (nkfb_score_dot) followed by the weighted sum of conjugate
    log-derivatives sum_s (u_s - <u>) conj O_k(x_s) / n (nkfb_weighted_score); with torch.distributed initialised the
    two sums are all-reduced, as NetKet's MPI means are.  Real-parameter models use Re S (NetKet's non-holomorphic
    real case).  The conjugate-gradient vectors live on the device in double precision whatever the state's
    precision."""

    def __init__(self, diag_shift=0.01, *, solver_tol=1e-5, maxiter=1000, **unused):
        self.diag_shift = float(diag_shift)
        self.solver_tol = float(solver_tol)
        self.maxiter = int(maxiter)
        self.info = None

    def qgt_matvec(self, vstate, v):
        """(S v) for a flat complex128 vector v (P,), using the state's current samples."""
        from ..parallel import allreduce_sum_
        h = vstate._h
        spec = vstate._spec()
        pk = vstate.samples_packed()
        flat = pk.view(-1, pk.shape[-1])
        real = not vstate.model.is_complex
        if real:
            v = v.real.to(torch.complex128)
        u = h.score_dot(spec, flat, vstate._ls, v.to(vstate._cdtype)).to(torch.complex128)
        tot = torch.stack([u.sum(), torch.tensor(float(u.numel()), dtype=torch.complex128, device=u.device)])
        tot_r = torch.view_as_real(tot).contiguous()
        allreduce_sum_(tot_r)
        n = float(tot_r[1, 0].item())
        mean = torch.complex(tot_r[0, 0], tot_r[0, 1]) / n
        acc = h.weighted_score(spec, flat, vstate._ls, (u - mean) / n)
        allreduce_sum_(acc)
        out = torch.view_as_complex(acc.view(-1, 2).contiguous())
        return out.real.to(torch.complex128) if real else out

    def __call__(self, vstate, grad, step=None):
        flat = grad if torch.is_tensor(grad) else vstate._flatten_tree(grad)
        b = flat.to(torch.complex128)
        shift = self.diag_shift

        def A(v):
            return self.qgt_matvec(vstate, v) + shift * v

        x = torch.zeros_like(b)
        r = b.clone()
        p = r.clone()
        rs = torch.vdot(r, r).real
        b2 = float(rs.item())
        it = 0
        if b2 > 0.0:
            for it in range(1, self.maxiter + 1):
                Ap = A(p)
                alpha = rs / torch.vdot(p, Ap).real
                x = x + alpha * p
                r = r - alpha * Ap
                rs_new = torch.vdot(r, r).real
                if float(rs_new.item()) <= (self.solver_tol ** 2) * b2:
                    rs = rs_new
                    break
                p = r + (rs_new / rs) * p
                rs = rs_new
        self.info = {"iterations": it, "residual": (float(rs.item()) / b2) ** 0.5 if b2 > 0 else 0.0}
        return vstate._tree(x.to(vstate._cdtype))

    def __repr__(self):
        return f"SR(diag_shift={self.diag_shift}, solver=cg(tol={self.solver_tol}), qgt=QGTOnTheFly)"
