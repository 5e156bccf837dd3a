"""Drop-in for the reference's LTS / PDRA stage render model ``app.fine.model.ESRNeRF``
(app/fine/model/esrnerf.py): same constructor arguments, ``state_dict`` keys / shapes and
``forward(**batch) -> Dict[str, Tensor]`` contract (``lts.py:331-333``, ``pdra.py:378-380``), with the per-sample
work in the sm_100a kernels of libesr_b200.so:

* primary rays: march + MaskCache + SDF tap (``differentiable_grid_sample`` arithmetic), NeuS alpha + transmittance
  scan + compactions, the analytic SDF gradient (``sample_sdf_expgrad``) with its backward into the grid, feature
  encode, the off / emo / emission / BRDF nets on tcgen05, tone mapper, compositing;
* the light-transport segment (esrnerf.py:487-679): hemisphere directions at ``num_ltspts`` shaded samples, point
  radiance for two view directions, ``num_ltspts x num_2ndrays`` secondary rays through the same march / scan /
  encode / MLP / composite kernels (near = ``lts_near``), SG environment map, Disney BRDF, Monte-Carlo mean;
* the two eps-jitter branches (esrnerf.py:807-830).

Random draws: the reference makes four per step (``np.random.choice``, ``torch.randn`` x3).  ``self.draws`` (an
object with ``choice(n, k)`` and ``randn(*shape)`` returning CPU tensors) overrides them — the parity tests feed the
oracle's numbers; by default numpy's host RNG picks the points (as in the reference) and the Gaussians are drawn on
the device.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import fused, pbr
from .modules import (BRDFNet, DenseGrid, EmissionNet, SphericalGaussian, cfg_get, flat_mlp_params_padded, host_geometry)
from .voxurff import VoxurfF


class ESRNeRF(VoxurfF):
    def __init__(self, cfg, near: float, far: float, xyz_min: torch.Tensor, xyz_max: torch.Tensor,
                 mask_xyz_min: torch.Tensor, mask_xyz_max: torch.Tensor, mask_alpha_init: float,
                 mask_density: torch.Tensor, s_val: float, num_voxles: int):
        self.pdra_mode = False
        self.draws = None
        # "numpy": np.random.choice on the host exactly as the reference (esrnerf.py:792; O(M3) host work per step);
        # "device": torch.randperm on the GPU — same distribution, different stream, no host round trip
        self.lts_sampler = "numpy"
        # sdf / off / emo grids, radiance nets and tone mapper are constructed exactly as in VoxurfF
        # (esrnerf.py:101-172 == voxurff.py:79-130), in the same order (same seeded initialisation)
        super().__init__(cfg, near, far, xyz_min, xyz_max, mask_xyz_min, mask_xyz_max, mask_alpha_init, mask_density,
                         s_val, num_voxles)
        m = "app.model."
        self.brdfnet_width = cfg_get(cfg, m + "brdfnet_width")
        self.brdfnet_depth = cfg_get(cfg, m + "brdfnet_depth")
        self.env_sg = cfg_get(cfg, m + "env_sg")
        self.env_activation = cfg_get(cfg, m + "env_activation")
        self.ray_sampling = cfg_get(cfg, m + "ray_sampling")
        self.num_2ndrays = cfg_get(cfg, m + "num_2ndrays")
        self.num_ltspts = cfg_get(cfg, m + "num_ltspts")
        self.lts_near = cfg_get(cfg, m + "lts_near")
        if not (self.brdfnet_width <= 192 and self.brdfnet_depth == 4):
            raise NotImplementedError("libesr_b200 instantiates the shipped LTS shape (cfg/app/lts.yaml:25-43): brdfnet "
                                      "<=192 x 4")
        # esrnerf.py:188-192 (the reference's `match` has no default arm: an unknown name fails at the first use)
        if str(self.ray_sampling).lower() in ("random", "rand"):
            self.fib_sampling = False
        elif str(self.ray_sampling).lower() in ("fib", "fibo", "fibonacci"):
            self.fib_sampling = True
        else:
            raise ValueError(f"ray_sampling {self.ray_sampling!r}: expected random | fib (esrnerf.py:188-192)")
        # esrnerf.py:174-195
        self.brdf = DenseGrid(self.color_dim, self.world_size, self.xyz_min, self.xyz_max)
        dim0 = (3 + 3 * self.posbase_pe * 2) + self.color_dim + len(self.grad_feat) * 9 + 1
        self.emitnet = EmissionNet(dim0, self.brdfnet_width, self.brdfnet_depth)
        self.brdfnet = BRDFNet(dim0, self.brdfnet_width, self.brdfnet_depth, self.brdf)
        self.envmap = SphericalGaussian(self.env_sg, self.env_activation)
        self.to(self.device)
        self.train()

    # ------------------------------------------------------------------------------------------
    def train(self, mode=True, finetune=False):
        """esrnerf.py:218-239"""
        if mode and not finetune:
            self.forward = self.forward_training
            if hasattr(self, "emit_color"):
                del self.emit_color
        elif mode:
            # scene-editing finetune (pdra.py:1079-1088): the emission net keeps reading a frozen copy of emo_color
            self.forward = self.forward_finetune
            self.emit_color = DenseGrid(self.color_dim, self.world_size, self.xyz_min, self.xyz_max).to(self.device)
            self.emit_color.load_state_dict(self.emo_color.state_dict())
            for p in self.emit_color.parameters():
                p.requires_grad_(False)
        else:
            self.forward = self.forward_evaluate
            if hasattr(self, "emo_color") and not hasattr(self, "emit_color"):
                self.emit_color = self.emo_color
        return torch.nn.Module.train(self, mode)

    @torch.no_grad()
    def scale_volume_grid(self, num_voxels):
        super().scale_volume_grid(num_voxels)
        self.brdf.scale_volume_grid(self.world_size)

    def _pbr_scene(self, near: float, manual: bool):
        g = self.sdf.grid.shape
        md = self.mask_cache.density.shape
        h = host_geometry(self, self.stepsize)
        return fused.make_scene(h["xyz_min"], h["xyz_max"], g[2:], h["mask_xyz_min"], h["mask_xyz_max"], md[2:], near,
                                1e9, h["stepdist"], h["voxel_size"], self.mask_cache.act_shift, self.maskcache_thres,
                                self.fastcolor_thres, float(self.s_val), fd_eps=1e-12, sdf_tap_manual=manual)

    def _flats(self):
        return (self._flat("off"), self._flat("emo"), flat_mlp_params_padded(self.emitnet.layers(), "emit"),
                flat_mlp_params_padded(self.brdfnet.layers(), "brdf"))

    # the reference's random draws (see module docstring)
    def _choice(self, n: int, k: int, dev) -> torch.Tensor:
        if self.draws is not None:
            return self.draws.choice(n, k).to(dev)
        if self.lts_sampler == "device":
            return torch.randperm(n, device=dev)[:k]
        return torch.from_numpy(np.random.choice(n, k, replace=False)).to(dev)

    def _randn(self, *shape, dev) -> torch.Tensor:
        if self.draws is not None:
            return self.draws.randn(*shape).to(dev)
        return torch.randn(*shape, device=dev)

    def _scatter(self, normal, number) -> torch.Tensor:
        """self.scattering(normal, number) of esrnerf.py:188-192: `number` directions in the hemisphere of each normal —
        normalised Gaussian draws (pbr/functions.py:10-18) or the fixed Fibonacci spiral (pbr/functions.py:21-32)"""
        if not normal.is_cuda:     # (host tensors: the torch restatement, used by CPU-side unit tests of the glue only)
            if self.fib_sampling:
                return pbr.diffuse_scattering_fib(normal, number)
            return pbr.diffuse_scattering(normal, self._randn(normal.shape[0], number, 3, dev=normal.device))
        if self.fib_sampling:
            return fused.lts_scatter_dirs(normal, number, table=pbr.fibonacci_hemisphere(number).to(normal.device))
        return fused.lts_scatter_dirs(normal, number, noise=self._randn(normal.shape[0], number, 3, dev=normal.device))

    def _env_radiance(self, dirs, last, add):
        """add + envmap(dirs) * last[:, None] (esrnerf.py:560-566; SphericalGaussian.forward, pbr/module.py:133-143): the
        fused kernel for the activations it implements, the module's own torch forward for any other name the reference's
        getattr lookup accepts"""
        env = self.envmap
        act_id = fused.SG_ACT_IDS.get(getattr(env.activation, "__name__", ""))
        if act_id is None or not dirs.is_cuda or env.mus.shape[0] > 64:
            return add + env(dirs) * last.unsqueeze(-1)
        lobes = F.normalize(env.lobes, dim=-1)
        lambdas = torch.abs(env.lambdas).reshape(-1)
        return fused.SgEnvmap.apply(dirs, env.mus, lambdas, lobes, act_id, last, add)

    def _shade(self, sc, pos, use, flats, emo_grid=None, sdf_grid=None):
        grids = (self.sdf.grid if sdf_grid is None else sdf_grid, self.off_color.grid,
                 self.emo_color.grid if emo_grid is None else emo_grid, self.brdf.grid if use[3] else None)
        fl = [f if u else None for f, u in zip(flats, use)]
        return fused.ShadePBR.apply(*grids, *fl, sc, pos, use, self._precision())

    # ------------------------------------------------------------------------------------------
    def _secondary(self, flats, rays_o2, d_flat, use=(True, True, False, False)):
        """The render chain over the LTS secondary rays (esrnerf.py:576-652 / 895-981): march from `lts_near`, scan,
        both radiance nets on every shaded sample, composite -> (sum w*off, sum w*emo, T_last) per secondary ray."""
        sc2 = self._pbr_scene(self.lts_near, False)
        st2 = fused.march(sc2, rays_o2, d_flat, None, self.mask_cache.density, self.sdf.grid.detach())
        sdf_g = self.sdf.grid if torch.is_grad_enabled() else self.sdf.grid.detach()
        if self.neus_alpha == "grad":      # esrnerf.py:346-361: the secondary rays' directions are their view directions
            st2.viewdirs = d_flat
        hw2, last2 = fused.AlphaScan.apply(sdf_g, sc2, rays_o2, d_flat, st2, None)
        pos2 = fused.SamplePos(st2.m3, d_flat, st2.h_sdf, rays_o2, d_flat, st2.h_ray, st2.h_step)
        lo, le, _, _ = self._shade(sc2, pos2, use, flats)
        if use[0]:
            off_m, emo_m = fused.Composite.apply(hw2, lo, le, st2)
        else:
            off_m, emo_m = None, fused.Composite.apply(hw2, le, None, st2)[0]
        return off_m, emo_m, last2, st2, hw2

    def _light_transport_segment(self, flats, pts, viewdirs, normal, sdf, base, rough, metal, emission, umask):
        """esrnerf.py:487-679"""
        dev = pts.device
        n2, P = self.num_2ndrays, pts.shape[0]
        dirs = self._scatter(normal, n2 + 1)
        v_rand = -dirs[:, -1]
        dirs = dirs[:, :-1]
        sc_pts = self._pbr_scene(self.near, False)
        # radiance leaving the points towards the camera and towards one random direction (esrnerf.py:499-547)
        pos = fused.SamplePos(2 * P, torch.cat([viewdirs, v_rand], 0).contiguous(), sdf.repeat(2).contiguous(),
                              pts=pts.repeat(2, 1).contiguous())
        off, emo, _, _ = self._shade(sc_pts, pos, (True, True, False, False), flats)

        def ex(t, c):
            return t.view(-1, 1, c).expand(P, n2, c).flatten(0, 1)

        d_flat = dirs.flatten(0, 1).contiguous()
        # incoming radiance: the secondary rays go through the whole render chain (esrnerf.py:576-652)
        off_m, emo_m, last2, st2, hw2 = self._secondary(flats, ex(pts, 3).contiguous(), d_flat)
        rad_off = self._env_radiance(d_flat, last2, off_m)       # off_m + envmap(d) * T_last  (esrnerf.py:560-566)
        # Disney reflectance x marched radiance, Monte-Carlo mean over the directions, both outgoing directions, and the
        # emission / PDRA mix of esrnerf.py:668-677: one kernel (esr_lts_accumulate) instead of the reference's
        # elementwise swarm (esrnerf.py:556-574, 654-677)
        off_hat, emo_hat = fused.LtsAccumulate.apply(base, rough.reshape(-1), metal.reshape(-1), rad_off, emo_m, normal,
                                                     -viewdirs, -v_rand, d_flat, n2, emission, umask, self.pdra_mode)
        if self.keep_streams:
            self.last_streams["lts"] = dict(streams=st2, h_w=hw2.detach())
        return dict(off=off, emo=emo, off_hat=off_hat, emo_hat=emo_hat)

    def forward_training(self, **kwargs) -> Dict[str, torch.Tensor]:
        """esrnerf.py:681-851"""
        rays_o = kwargs["rays_o"].contiguous().float()
        rays_d = kwargs["rays_d"].contiguous().float()
        viewdirs = kwargs["viewdirs"].contiguous().float()
        em_modes = kwargs["em_modes"].long().contiguous()
        uncert_masks = kwargs["uncert_masks"]
        self.s_val = kwargs["s_val"]
        normal_eps, emit_eps = kwargs["normal_eps"], kwargs["emit_eps"]
        dev = rays_o.device
        with torch.cuda.device(dev):
            flats = self._flats()
            flat_tone = self._flat("tone")
            for g in (self.sdf, self.off_color, self.emo_color, self.brdf):
                g.ensure_layout()
            sc = self._pbr_scene(self.near, True)
            # the uncertain-ray count rides along on the host read that sizes the M1 stream: the two boolean-mask
            # outputs below then need no synchronisation of their own
            um = uncert_masks.bool()
            s, n_uncert = fused.march(sc, rays_o, rays_d, None, self.mask_cache.density, self.sdf.grid.detach(),
                                      also_read=um.sum(dtype=torch.int32))
            um_order = torch.argsort((~um).to(torch.uint8), stable=True)     # uncertain rays first, original order kept
            if self.neus_alpha == "grad":
                s.viewdirs = viewdirs
            h_w, last = fused.AlphaScan.apply(self.sdf.grid, sc, rays_o, rays_d, s, None)
            m3 = s.m3
            pts = fused.sample_points(sc, rays_o, rays_d, s.h_ray, s.h_step)
            exp_grad = fused.SdfExpGrad.apply(self.sdf.grid, sc, pts)
            pos = fused.SamplePos(m3, viewdirs, s.h_sdf, rays_o, rays_d, s.h_ray, s.h_step)
            lin_off, lin_emo, emit, brdf = self._shade(sc, pos, (True, True, True, True), flats)
            # esrnerf.py:751-757: emo on the emission-on rays + off on all of them, no stop-gradient
            rgb, lin = fused.CombineTonemap.apply(lin_off, lin_emo, flat_tone, s.h_ray, em_modes, False, True,
                                                  self._precision())
            rgb_m, lin_m = fused.Composite.apply(h_w, rgb, lin, s)
            emit_m, _ = fused.Composite.apply(h_w, emit, None, s)
            if self.keep_streams:
                self.last_streams = dict(streams=s, h_w=h_w.detach(), lin=lin.detach(), rgb=rgb.detach(), pts=pts)

            normal = F.normalize(exp_grad.detach(), dim=-1)
            idx = self._choice(m3, min(self.num_ltspts, m3), dev)
            ray_l = s.h_ray.long()[idx]
            base, rough, metal = brdf.split([3, 1, 1], -1)
            lts = self._light_transport_segment(flats, pts[idx], viewdirs[ray_l], normal[idx], s.h_sdf[idx], base[idx],
                                                rough[idx], metal[idx], emit[idx], uncert_masks[ray_l])
            # eps branches (esrnerf.py:807-830)
            exp_grad_eps = fused.SdfExpGrad.apply(self.sdf.grid, sc,
                                                  pts + self._randn(m3, 3, dev=dev) * normal_eps)
            pts_e = (pts + self._randn(m3, 3, dev=dev) * emit_eps).contiguous()
            sdf_e = fused.sdf_tap_points(sc, self.sdf.grid.detach(), pts_e)
            pos_e = fused.SamplePos(m3, pts_e, sdf_e, pts=pts_e)          # view columns are not inputs of these nets
            _, _, emit_e, brdf_e = self._shade(sc, pos_e, (False, False, True, True), flats)
        return {
            "etc/alphainv_cum": last,
            "etc/white_bg": last[..., None],
            "srgb/rgb": rgb_m,
            "lin/rgb": lin_m,
            "lin/pbr/off": lts["off"],
            "lin/pbr/off_hat": lts["off_hat"],
            "lin/pbr/emo": lts["emo"],
            "lin/pbr/emo_hat": lts["emo_hat"],
            "etc/emit_uncert": emit_m[um_order[:n_uncert]],        # == emit_m[uncert_masks]  (esrnerf.py:840)
            "etc/emit_cert": emit_m[um_order[n_uncert:]],          # == emit_m[~uncert_masks] (esrnerf.py:841)
            "etc/normal": exp_grad,
            "etc/normal_eps": exp_grad_eps,
            "etc/emit": emit,
            "etc/emit_eps": emit_e,
            "etc/brdf": brdf,
            "etc/brdf_eps": brdf_e,
        }

    @torch.no_grad()
    def forward_finetune(self, **kwargs) -> Dict[str, torch.Tensor]:
        """esrnerf.py:241-484 — scene-editing finetune target.  Everything runs without autograd except
        emo_rgbnet(emo_color(x)) at the LTS points: 'lin/pbr/emo' carries gradient to emo_rgbnet /
        emo_color only, 'lin/pbr/emo_hat' = edited emission + reflected emission is a constant target."""
        rays_o, rays_d, viewdirs = self._rays(kwargs)
        em_modes, em_int, em_col = kwargs["em_modes"], kwargs["em_intensities"], kwargs["em_colors"]
        dev = rays_o.device
        n2 = self.num_2ndrays
        with torch.cuda.device(dev):
            sc, s, _, _, _ = self._eval_stream(rays_o, rays_d, False, viewdirs)
            idx = self._choice(s.m3, min(self.num_ltspts, s.m3), dev)
            P = idx.shape[0]
            if P == 0:
                z = torch.zeros(0, 3, device=dev)
                return {"lin/pbr/emo": z, "lin/pbr/emo_hat": z}
            ray = s.h_ray.long()[idx]
            pts = fused.sample_points(sc, rays_o, rays_d, s.h_ray[idx].contiguous(), s.h_step[idx].contiguous())
            vdir = viewdirs[ray]
            sc_pts = self._pbr_scene(self.near, False)
            sdf, exp_grad = fused.sdf_expgrad_points(sc_pts, self.sdf.grid.detach(), pts)
            normal = F.normalize(exp_grad, dim=-1)
            dirs = self._scatter(normal, n2 + 1)
            v_rand = -dirs[:, -1]
            dirs = dirs[:, :-1]
            flats_ng = [f.detach() for f in self._flats()]
            pos2 = fused.SamplePos(2 * P, torch.cat([vdir, v_rand], 0).contiguous(), sdf.repeat(2).contiguous(),
                                   pts=pts.repeat(2, 1).contiguous())
            with torch.enable_grad():   # the features are constants here: only emo_rgbnet / emo_color receive gradient
                flat_emo = self._flat("emo")
                _, emo, _, _ = self._shade(sc_pts, pos2, (False, True, False, False), [None, flat_emo, None, None],
                                           sdf_grid=self.sdf.grid.detach())
            pos1 = fused.SamplePos(P, vdir.contiguous(), sdf, pts=pts)
            _, _, emit, brdf = self._shade(sc_pts, pos1, (False, False, True, True), flats_ng,
                                           emo_grid=self.emit_color.grid.detach())

            def ex(t, c):
                return t.view(-1, 1, c).expand(P, n2, c).flatten(0, 1)

            d_flat = dirs.flatten(0, 1).contiguous()
            _, emo_m, _, _, _ = self._secondary(flats_ng, ex(pts, 3).contiguous(), d_flat, (False, True, False, False))
            emit = pbr.edit_emission(emit, em_modes[ray], em_int[ray], em_col[ray])
            _, reflect = fused.LtsAccumulate.apply(brdf[:, :3], brdf[:, 3], brdf[:, 4], None, emo_m, normal, -vdir, -v_rand,
                                                   d_flat, n2)
        return {"lin/pbr/emo": emo, "lin/pbr/emo_hat": emit.repeat(2, 1) + reflect}

    @torch.no_grad()
    def render_envmap(self, H: int, W: int) -> torch.Tensor:
        """esrnerf.py:1668-1690: the SG environment map on an equirectangular H x W lattice -> [H,W,3]"""
        dev = self.envmap.mus.device
        phi, theta = torch.meshgrid(torch.linspace(0.0, np.pi, H, device=dev),
                                    torch.linspace(1.0 * np.pi, -1.0 * np.pi, W, device=dev), indexing="ij")
        dirs = torch.stack([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], -1)
        return self.envmap(dirs.view(-1, 3)).view(H, W, 3)

    # ------------------------------------------------------------------------------------------
    # inference entry points
    # ------------------------------------------------------------------------------------------
    def _eval_stream(self, rays_o, rays_d, manual: bool, viewdirs=None):
        """shared head of forward_evaluate / eval_emit / eval_esp: packed shaded stream without autograd state.
        Returns (scene, streams, h_w, last, degenerate) — `degenerate` flags the reference's `.squeeze()` quirk
        (SURVEY.md Q7): exactly one sample passing the alpha filter makes 0-dim tensors there and zero images."""
        sc = self._pbr_scene(self.near, manual)
        for g in (self.sdf, self.off_color, self.emo_color, self.brdf):
            g.ensure_layout()
        s = fused.march(sc, rays_o, rays_d, None, self.mask_cache.density, self.sdf.grid.detach())
        if self.neus_alpha == "grad":
            s.viewdirs = viewdirs
        h_w, last = fused.AlphaScan.apply(self.sdf.grid.detach(), sc, rays_o, rays_d, s, None)
        degenerate = s.m3 <= 1 and s.m1 > 0 and int((s.s_alpha > self.fastcolor_thres).sum()) == 1
        return sc, s, h_w, last, degenerate

    @staticmethod
    def _rays(kwargs):
        return (kwargs["rays_o"].contiguous().float(), kwargs["rays_d"].contiguous().float(),
                kwargs["viewdirs"].contiguous().float())

    @torch.no_grad()
    def eval_esp(self, **kwargs) -> torch.Tensor:
        """esrnerf.py:1360-1407: expected surface point sum_ray w * ray_pts -> [N,3]"""
        rays_o, rays_d, viewdirs = self._rays(kwargs)
        with torch.cuda.device(rays_o.device):
            sc, s, h_w, _, degenerate = self._eval_stream(rays_o, rays_d, False, viewdirs)
            if degenerate:
                return torch.zeros_like(rays_o)
            pts = fused.sample_points(sc, rays_o, rays_d, s.h_ray, s.h_step)
            return fused.composite_infer(h_w, pts, pts, s)[0]

    @torch.no_grad()
    def eval_emit(self, **kwargs) -> torch.Tensor:
        """esrnerf.py:1299-1358: composite of the emission net -> [N,3]"""
        rays_o, rays_d, viewdirs = self._rays(kwargs)
        with torch.cuda.device(rays_o.device):
            sc, s, h_w, _, degenerate = self._eval_stream(rays_o, rays_d, False, viewdirs)
            if degenerate:
                return torch.zeros_like(rays_o)
            pos = fused.SamplePos(s.m3, viewdirs, s.h_sdf, rays_o, rays_d, s.h_ray, s.h_step)
            flats = [f.detach() for f in self._flats()]
            _, _, emit, _ = self._shade(sc, pos, (False, False, True, False), flats)
            return fused.composite_infer(h_w, emit, emit, s)[0]

    def _lts_eval(self, flats, pts, viewdirs, normal, base, rough, metal, emit):
        """esrnerf.py:854-1001 (one chunk of shaded samples): environment / emission light decomposed into direct
        and indirect parts by Monte-Carlo integration over `num_2ndrays` hemisphere directions per sample."""
        dev = pts.device
        n2, P = self.num_2ndrays, pts.shape[0]
        dirs = self._scatter(normal, n2)

        def ex(t, c):
            return t.view(-1, 1, c).expand(P, n2, c).flatten(0, 1)

        d_flat = dirs.flatten(0, 1).contiguous()
        R = pbr.disney_reflection(ex(base, 3), ex(rough, 1), ex(metal, 1), ex(normal, 3), d_flat, -ex(viewdirs, 3))
        off_m, emo_m, last2, _, _ = self._secondary(flats, ex(pts, 3).contiguous(), d_flat)
        env = self.envmap(d_flat) * last2.unsqueeze(-1)
        out = {"lin/env_dir": (env * R).view(-1, n2, 3).mean(-2), "lin/env_indir": (off_m * R).view(-1, n2, 3).mean(-2)}
        out["lin/env_effects"] = out["lin/env_dir"] + out["lin/env_indir"]
        out["lin/emit_(in)dir"] = (emo_m * R).view(-1, n2, 3).mean(-2)
        out["lin/emit_effects"] = emit + out["lin/emit_(in)dir"]
        return out

    PBR_KEYS = ("lin/env_dir", "lin/env_indir", "lin/env_effects", "lin/emit_(in)dir", "lin/emit_effects")

    @torch.no_grad()
    def forward_evaluate(self, **kwargs) -> Dict[str, torch.Tensor]:
        """esrnerf.py:853-1297 — inference: the 12 VoxurfF maps + emission / BRDF maps and, with `render_pbr`, the
        light-transport decomposition of every shaded sample (in chunks of `chunk_sz` samples)."""
        rays_o, rays_d, viewdirs = self._rays(kwargs)
        em_modes = kwargs["em_modes"]
        render_pbr, chunk_sz = kwargs["render_pbr"], kwargs["chunk_sz"]
        dev = rays_o.device
        pos_rt = kwargs["pos_rt"].to(dev).float()
        assert getattr(self, "emit_color", self.emo_color) is self.emo_color      # eval aliases it (esrnerf.py:236-238)
        with torch.cuda.device(dev):
            sc, s, h_w, last, degenerate = self._eval_stream(rays_o, rays_d, True, viewdirs)
            if degenerate:
                z3 = torch.zeros_like(rays_o)
                depth = z3[..., 0]
                out = {"etc/depth": depth, "etc/disp": 1 / (depth + self.far), "etc/normal": z3,
                       "etc/white_bg": torch.ones_like(z3[..., :1])}
                for k in ("srgb/off_rgb", "lin/off_rgb", "srgb/on_rgb", "lin/on_rgb", "srgb/emo_rgb", "lin/emo_rgb",
                          "srgb/rgb", "lin/rgb", "lin/emit", "lin/basecolor"):
                    out[k] = z3
                out["lin/roughness"] = out["lin/metallic"] = depth
                if render_pbr:
                    out.update({k: z3 for k in self.PBR_KEYS})
                return out
            m3 = s.m3
            flats = [f.detach() for f in self._flats()]
            pos = fused.SamplePos(m3, viewdirs, s.h_sdf, rays_o, rays_d, s.h_ray, s.h_step)
            lin_off, lin_emo, emit, brdf = self._shade(sc, pos, (True, True, True, True), flats)
            lin_on = lin_off + lin_emo
            srgb = fused.tonemap_infer(torch.cat([lin_off, lin_on, lin_emo], 0), self._flat("tone").detach())
            off_rgb, on_rgb, emo_rgb = srgb[:m3], srgb[m3: 2 * m3], srgb[2 * m3:]
            grad = fused.sdf_fd_gradient(sc, rays_o, rays_d, self.sdf.grid.detach(), s)
            normal = (F.normalize(grad, dim=-1) @ pos_rt * self.normal_flipper.to(dev) + 1.0) / 2.0
            aux = torch.zeros(m3, 3, device=dev)          # (step * dist, roughness, metallic) share one composite
            aux[:, 0] = s.h_step.float() * host_geometry(self, self.stepsize)["stepdist"]
            aux[:, 1:] = brdf[:, 3:5]
            base = brdf[:, :3].contiguous()
            off_m, lin_off_m = fused.composite_infer(h_w, off_rgb, lin_off, s)
            on_m, lin_on_m = fused.composite_infer(h_w, on_rgb, lin_on, s)
            emo_m, lin_emo_m = fused.composite_infer(h_w, emo_rgb, lin_emo, s)
            normal_m, aux_m = fused.composite_infer(h_w, normal, aux, s)
            emit_m, base_m = fused.composite_infer(h_w, emit, base, s)
            lts = {}
            if render_pbr and m3:
                pts = fused.sample_points(sc, rays_o, rays_d, s.h_ray, s.h_step)
                exp_grad = fused.SdfExpGrad.apply(self.sdf.grid.detach(), sc, pts)
                nrm = F.normalize(exp_grad, dim=-1)
                vdir = viewdirs[s.h_ray.long()]
                parts = {k: [] for k in self.PBR_KEYS}
                for idx in torch.arange(m3, device=dev).split(int(chunk_sz)):
                    ret = self._lts_eval(flats, pts[idx], vdir[idx], nrm[idx], base[idx], brdf[idx, 3:4], brdf[idx, 4:5],
                                         emit[idx])
                    for k in self.PBR_KEYS:
                        parts[k].append(ret[k])
                cat = [torch.cat(parts[k], 0) for k in self.PBR_KEYS]
                m01 = fused.composite_infer(h_w, cat[0], cat[1], s)
                m23 = fused.composite_infer(h_w, cat[2], cat[3], s)
                m4 = fused.composite_infer(h_w, cat[4], cat[4], s)
                lts = dict(zip(self.PBR_KEYS, (*m01, *m23, m4[0])))
            elif render_pbr:
                lts = {k: torch.zeros_like(rays_o) for k in self.PBR_KEYS}
        depth = aux_m[:, 0].contiguous()
        em = int(em_modes) if not torch.is_tensor(em_modes) else int(em_modes.item())
        rgb_m, lin_rgb_m = (off_m, lin_off_m) if em == 0 else (on_m, lin_on_m)
        if self.keep_streams:
            self.last_streams = dict(streams=s, h_w=h_w)
        return {"etc/depth": depth, "etc/disp": 1 / (depth + last * self.far), "etc/normal": normal_m,
                "etc/white_bg": last.unsqueeze(-1), "srgb/off_rgb": off_m, "lin/off_rgb": lin_off_m, "srgb/on_rgb": on_m,
                "lin/on_rgb": lin_on_m, "srgb/emo_rgb": emo_m, "lin/emo_rgb": lin_emo_m, "srgb/rgb": rgb_m,
                "lin/rgb": lin_rgb_m, "lin/emit": emit_m, "lin/basecolor": base_m,
                "lin/roughness": aux_m[:, 1].contiguous(), "lin/metallic": aux_m[:, 2].contiguous(), **lts}
