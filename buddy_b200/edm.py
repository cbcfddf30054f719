"""EDM parameterisation — drop-in for `diff_params.edm.EDM` (reference diff_params/edm.py:7-96,
diff_params/shared.py:98-120).  Scalar algebra only; the network call is the CUDA engine."""
import torch


class EDM:
    def __init__(self, type="ve_karras", sde_hp=None):
        self.type = type
        self.sde_hp = sde_hp
        get = (lambda k: sde_hp[k]) if isinstance(sde_hp, dict) else (lambda k: getattr(sde_hp, k))
        self.sigma_data = get("sigma_data")
        self.sigma_min = get("sigma_min")
        self.sigma_max = get("sigma_max")
        self.rho = get("rho")

    def cskip(self, sigma):
        return self.sigma_data ** 2 * (sigma ** 2 + self.sigma_data ** 2) ** -1

    def cout(self, sigma):
        return sigma * self.sigma_data * (self.sigma_data ** 2 + sigma ** 2) ** (-0.5)

    def cin(self, sigma):
        return (self.sigma_data ** 2 + sigma ** 2) ** (-0.5)

    def cnoise(self, sigma):
        return (1 / 4) * torch.log(sigma)

    def _mean(self, x, t):
        return x

    def _std(self, t):
        return t

    def Tweedie2score(self, tweedie, xt, t, *args, **kwargs):
        return (tweedie - self._mean(xt, t)) / self._std(t) ** 2

    def score2Tweedie(self, score, xt, t, *args, **kwargs):
        return self._std(t) ** 2 * score + self._mean(xt, t)

    def _ode_integrand(self, x, t, score):
        return -t * score

    def denoiser(self, xn, net, t, *args, **kwargs):
        """cskip*xn + cout*net(cin*xn, cnoise)  (shared.py:98-120); xn (B,1,T)."""
        sigma = self._std(torch.as_tensor(t, dtype=torch.float32, device=xn.device)).unsqueeze(-1)
        sigma = sigma.view(*sigma.size(), *(1,) * (xn.ndim - sigma.ndim))
        cnoise = self.cnoise(sigma.squeeze())
        cnoise = cnoise.repeat(xn.shape[0], ) if cnoise.dim() == 0 else cnoise.view(xn.shape[0], )
        return self.cskip(sigma) * xn + self.cout(sigma) * net(self.cin(sigma) * xn, cnoise)
