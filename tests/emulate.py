"""CPU emulation of the kernel-side arithmetic on the PACKED tables (what libswc uploads), used to
validate the host packing and the closed-form restatements (tap tables, frame stacking order,
deconvolution parity split, anti-aliased snake stencil, interleaved head + inverse-DFT operand,
overlap-add) against the oracle without a GPU.  It mirrors simwhisper_codec_b200/csrc/pipeline.cu.
Test-only code."""
import torch
import torch.nn.functional as F


def tap_gemm(A, W, taps, tap_k, m_rows, bias=None):
    """A (nb,R,C), W (N, len(taps)*tap_k); out[b,m,:] = sum_i A[b, m+row_i, col_i:col_i+tap_k] @ W_i^T."""
    nb, R, _ = A.shape
    out = torch.zeros(nb, m_rows, W.shape[0], dtype=A.dtype)
    for i, (tr, tc) in enumerate(taps):
        lo, hi = max(0, -tr), min(m_rows, R - tr)
        if hi > lo:
            out[:, lo:hi] += A[:, lo + tr:hi + tr, tc:tc + tap_k] @ W[:, i * tap_k:(i + 1) * tap_k].T
    if bias is not None:
        out = out + bias
    return out


class Emu:
    def __init__(self, nat, dtype=torch.float64):
        self.nat = nat
        self.dt = dtype

    def t(self, name, *shape):
        x = self.nat.packed(name).to(self.dt)
        return x.view(*shape) if shape else x

    def ln(self, x, g, b, eps):
        return F.layer_norm(x, (x.shape[-1],), self.t(g), self.t(b), eps)

    def layer(self, p, h, lens):
        nb, T, D = h.shape
        x = self.ln(h, p + ".ln1.g", p + ".ln1.b", 1e-5)
        qkv = x @ self.t(p + ".qkv.w", 3 * D, D).T + self.t(p + ".qkv.b")
        q, k, v = (u.view(nb, T, 12, 64).transpose(1, 2) for u in qkv.split(D, dim=-1))
        s = q @ k.transpose(-1, -2)
        key_ok = torch.arange(T)[None, :] < lens[:, None]
        s = s.masked_fill(~key_ok[:, None, None, :], float("-inf"))
        o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(nb, T, D)
        h = h + o @ self.t(p + ".out.w", D, D).T + self.t(p + ".out.b")
        x = self.ln(h, p + ".ln2.g", p + ".ln2.b", 1e-5)
        f = F.gelu(x @ self.t(p + ".fc1.w", 3072, D).T + self.t(p + ".fc1.b"))
        return h + f @ self.t(p + ".fc2.w", D, 3072).T + self.t(p + ".fc2.b")

    def encoder(self, mel_cf, mel_lens, n_layers=12):
        nb, _, Tm = mel_cf.shape
        T, D = (Tm + 1) // 2, 768
        mel_cl = torch.zeros(nb, Tm, 128, dtype=self.dt)
        mel_cl[:, :, :80] = mel_cf.transpose(1, 2)
        stem = torch.zeros(nb, 2 * T, D, dtype=self.dt)
        stem[:, :Tm] = tap_gemm(mel_cl, self.t("enc.conv1.w", D, 384), [(-1, 0), (0, 0), (1, 0)], 128, Tm, self.t("enc.conv1.b"))
        pair = stem.view(nb, T, 2 * D)
        h = tap_gemm(pair, self.t("enc.conv2.w", D, 3 * D), [(-1, D), (0, 0), (0, D)], D, T, self.t("enc.conv2.b"))
        lens = mel_lens // 2
        for i in range(n_layers):
            h = self.layer(f"enc.L{i}", h, lens)
        h = self.ln(h, "enc.ln.g", "enc.ln.b", 1e-5)
        keep = (torch.arange(T)[None, :] < lens[:, None])[..., None]
        h = torch.where(keep, h, torch.zeros((), dtype=self.dt))
        T4 = (T + 3) // 4 * 4
        out = torch.zeros(nb, T4, D, dtype=self.dt)
        out[:, :T] = h
        return out, lens

    def aa_snake(self, p, x):
        """closed form of kernels.cu::aa_snake_kernel on channel-last (nb,T,C)."""
        nb, T, C = x.shape
        fu, fd = self.t(p + ".fu"), self.t(p + ".fd")
        ea = torch.exp(self.t(p + ".alpha"))
        inv_b = 1.0 / (torch.exp(self.t(p + ".beta")) + 1e-9)
        m = torch.arange(2 * T)
        q, odd = m // 2, m % 2
        u = torch.zeros(nb, 2 * T, C, dtype=self.dt)
        for a in range(6):
            j = (q - 3 + odd + a).clamp(0, T - 1)
            tap = torch.where(odd == 1, fu[10 - 2 * a], fu[11 - 2 * a])
            u += x[:, j] * tap[None, :, None]
        u = 2.0 * u
        v = u + inv_b * torch.sin(u * ea) ** 2
        y = torch.zeros(nb, T, C, dtype=self.dt)
        t = torch.arange(T)
        for k in range(12):
            y += v[:, (2 * t + k - 5).clamp(0, 2 * T - 1)] * fd[k]
        return y

    def res_units(self, pfx, x):
        nb, Tc, H = x.shape
        for i, d in enumerate((1, 3, 9)):
            p = f"{pfx}.res{i}"
            a = self.aa_snake(p + ".act0", x)
            c = tap_gemm(a, self.t(p + ".conv7.w", H, 7 * H), [((k - 3) * d, 0) for k in range(7)], H, Tc, self.t(p + ".conv7.b"))
            e = self.aa_snake(p + ".act2", c)
            x = x + e @ self.t(p + ".conv1.w", H, H).T + self.t(p + ".conv1.b")
        return x

    def fsq(self, lat, lens):
        """lat (nb,Tc,32) -> dq (nb,Tc,32), codes (8,nb,Tc)"""
        c = self.nat.packed("fsq.const")
        scale, offset, shift, half = (c[i * 4:(i + 1) * 4] for i in range(4))
        base = torch.tensor([1, 8, 56, 336])
        nb, Tc, _ = lat.shape
        x = lat.to(torch.float32).view(nb, Tc, 8, 4)
        comp = scale * torch.tanh(x + shift) - offset
        r = torch.round(comp)
        dq = r / half
        idx = ((r + half).to(torch.int64) * base).sum(-1)
        keep = torch.arange(Tc)[None, :] < lens[:, None]
        return (dq * keep[..., None, None]).view(nb, Tc, 32), (idx * keep[..., None]).permute(2, 0, 1).to(torch.int32)

    def downsample(self, enc_cl, enc_lens):
        nb, T4, D = enc_cl.shape
        Tc, H = T4 // 4, 512
        x = enc_cl.reshape(nb, Tc, 4 * D) @ self.t("dn.in.w", H, 4 * D).T + self.t("dn.in.b")
        x = self.res_units("dn", x)
        lat = x @ self.t("dn.latent.w", 32, H).T + self.t("dn.latent.b")
        return lat, (enc_lens + 3) // 4

    def upsample(self, zq_cl):
        nb, Tc, _ = zq_cl.shape
        x = zq_cl.to(self.dt) @ self.t("up.from.w", 512, 32).T + self.t("up.from.b")
        x = self.res_units("up", x)
        h = x @ self.t("up.stacked.w", 3072, 512).T + self.t("up.stacked.b")
        return h.reshape(nb, 4 * Tc, 768)

    def decoder(self, h, lens, n_layers=12):
        nb, T, D = h.shape
        for i in range(n_layers):
            h = self.layer(f"dec.L{i}", h, lens)
        y = self.ln(h, "dec.ln.g", "dec.ln.b", 1e-5)
        keep = (torch.arange(T)[None, :] < lens[:, None])[..., None]
        y = torch.where(keep, y, torch.zeros((), dtype=self.dt))
        z = torch.zeros(nb, 2 * T, D, dtype=self.dt)
        b1 = self.t("dec.deconv1.b")
        z[:, 0::2] = tap_gemm(y, self.t("dec.deconv1e.w", D, 2 * D), [(-1, 0), (0, 0)], D, T, b1)
        z[:, 1::2] = tap_gemm(y, self.t("dec.deconv1o.w", D, D), [(0, 0)], D, T, b1)
        return tap_gemm(z, self.t("dec.deconv2.w", 128, 3 * D), [(0, 0), (-1, 0), (-2, 0)], D, 2 * T, self.t("dec.deconv2.b"))

    def vocos(self, mel_cl, n_blocks=24):
        nb, Tv, _ = mel_cl.shape
        V, I = 512, 4096
        e = tap_gemm(mel_cl, self.t("voc.embed.w", V, 7 * 128), [(k - 3, 0) for k in range(7)], 128, Tv, self.t("voc.embed.b"))
        x = self.ln(e, "voc.norm.g", "voc.norm.b", 1e-6)
        for i in range(n_blocks):
            p = f"voc.B{i}"
            w = self.t(p + ".dw.w", 7, V)
            y = torch.zeros_like(x) + self.t(p + ".dw.b")
            for k in range(7):
                lo, hi = max(0, 3 - k), min(Tv, Tv + 3 - k)
                y[:, lo:hi] += x[:, lo + k - 3:hi + k - 3] * w[k]
            y = self.ln(y, p + ".ln.g", p + ".ln.b", 1e-6)
            g = F.gelu(y @ self.t(p + ".pw1.w", I, V).T + self.t(p + ".pw1.b"))
            x = x + (g @ self.t(p + ".pw2.w", V, I).T + self.t(p + ".pw2.b")) * self.t(p + ".gamma")
        y = self.ln(x, "voc.final.g", "voc.final.b", 1e-6)
        hd = y @ self.t("voc.head.w", 704, V).T + self.t("voc.head.b")
        mag = torch.exp(hd[..., 0::2]).clamp(max=100.0)
        ph = hd[..., 1::2]
        S = torch.stack([mag * torch.cos(ph), mag * torch.sin(ph)], dim=-1).reshape(nb, Tv, 704)
        frames = S @ self.t("voc.idft.w", 640, 704).T
        w2 = self.t("voc.win_sq")
        L = 160 * Tv
        out = torch.zeros(nb, L, dtype=self.dt)
        sidx = torch.arange(L)
        p = sidx + 240
        env = torch.zeros(L, dtype=self.dt)
        for dt_ in range(4):
            t = p // 160 - dt_
            n = p - 160 * t
            ok = (t >= 0) & (t < Tv) & (n < 640)
            tt, nn_ = t.clamp(0, Tv - 1), n.clamp(0, 639)
            out += frames[:, tt, nn_] * ok
            env += w2[nn_] * ok
        return out / env

    def mel(self, wav, lens):
        """wav (nb, L) fp32 -> log-mel (nb,80,3000) via the packed windowed-DFT and filterbank operands."""
        nb = wav.shape[0]
        x = torch.zeros(nb, 480000, dtype=self.dt)
        for b in range(nb):
            n = min(int(lens[b]), 480000, wav.shape[1])
            x[b, :n] = wav[b, :n]
        xp = torch.cat([x[:, 1:201].flip(1), x, x[:, -201:-1].flip(1)], dim=1)      # reflect pad 200
        fr = xp.unfold(1, 400, 160)[:, :3000]                                        # (nb,3000,400)
        spec = fr @ self.t("mel.dft.w", 416, 400).T
        power = spec[..., 0::2] ** 2 + spec[..., 1::2] ** 2                           # (nb,3000,208)
        mel = power @ self.t("mel.fb.w", 80, 208).T
        lg = torch.log10(mel.clamp(min=1e-10))
        mx = lg.amax(dim=(1, 2), keepdim=True)
        return ((torch.maximum(lg, mx - 8.0) + 4.0) / 4.0).transpose(1, 2)
