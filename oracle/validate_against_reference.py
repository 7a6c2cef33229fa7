"""TEST INFRASTRUCTURE ONLY -- pin the oracle against the UNMODIFIED reference and write the golden fixtures.

Runs only in the build container (needs /root/reference).  It
 1. imports the reference (oracle/reference_import.py) on CPU,
 2. checks every function of oracle/aninerf_oracle.py against the reference function it restates,
    on seeded synthetic inputs (bit-equal on this machine: same ATen ops in the same order),
 3. writes tests/golden/*.npz: inputs + the REFERENCE's outputs for small cases, so that the oracle
    (and the CUDA path) can be re-checked anywhere the reference is not available,
 4. writes oracle/VALIDATION.md with the comparison table.

    python -m oracle.validate_against_reference
"""
from __future__ import annotations

import hashlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import aninerf_oracle as O          # noqa: E402
from oracle import reference_import            # noqa: E402
from animatable_nerf_b200 import synthetic     # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
rows = []


def report(name, ref_cite, ok, detail):
    rows.append((name, ref_cite, 'PASS' if ok else 'FAIL', detail))
    print(f'[{"PASS" if ok else "FAIL"}] {name}: {detail}')


def maxdiff(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    if a.shape != b.shape:
        return float('inf')
    return float((a - b).abs().max()) if a.numel() else 0.0


def biteq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = reference_import.load('configs/aninerf_313.yaml')
    cfg = ref.cfg
    os.makedirs(GOLDEN, exist_ok=True)

    # ---------------- stage 1: rays / near-far (numpy) ---------------------------------------
    frame = synthetic.make_frame(pose_seed=2, body_seed=1, voxel=0.025)
    K, R, T = synthetic.make_camera(frame, 1024, 1024)
    ro_r, rd_r = ref.dutils.get_rays(1024, 1024, K, R, T)
    ro_o, rd_o = O.get_rays(1024, 1024, K, R, T)
    report('get_rays 1024x1024', 'if_nerf_data_utils.py:64-89', biteq(rd_r, rd_o) and biteq(np.ascontiguousarray(ro_r), np.ascontiguousarray(ro_o)),
           'fp64 rays bit-equal')
    ref_out = ref.dutils.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'])
    ora_out = O.get_rays_within_bounds(1024, 1024, K, R, T, frame['wbounds'])
    ok = all(biteq(a, b) for a, b in zip(ref_out, ora_out))
    report('get_rays_within_bounds 1024x1024', 'if_nerf_data_utils.py:156-196,310-339', ok,
           f'{int(ref_out[4].sum())} of {ref_out[4].size} rays hit the box; ray_o/ray_d/near/far/mask bit-equal')
    A_ref = ref.dutils.get_rigid_transformation(frame['poses'].copy(), frame['joints'].copy(), synthetic.SMPL_PARENTS)
    A_ora = O.get_rigid_transformation(frame['poses'].copy(), frame['joints'].copy(), synthetic.SMPL_PARENTS)
    report('get_rigid_transformation', 'if_nerf_data_utils.py:414-458', biteq(A_ref, A_ora) and maxdiff(A_ref, frame['A']) < 1e-6,
           f'oracle bit-equal; host_geometry.bone_transforms max diff {maxdiff(A_ref, frame["A"]):.2e}')
    report('get_bounds', 'if_nerf_data_utils.py:566-579', biteq(ref.dutils.get_bounds(frame['wverts']), O.get_bounds(frame['wverts'])), 'bit-equal')

    # small golden for stage 1 (64x64 image, zoomed so that the box covers part of it)
    Ks, Rs, Ts = synthetic.make_camera(frame, 64, 64, focal=60.0)
    g_ro, g_rd = ref.dutils.get_rays(64, 64, Ks, Rs, Ts)
    g = ref.dutils.get_rays_within_bounds(64, 64, Ks, Rs, Ts, frame['wbounds'])
    np.savez_compressed(os.path.join(GOLDEN, 'stage1_rays_64.npz'), K=Ks, R=Rs, T=Ts, bounds=frame['wbounds'],
                        rays_o=g_ro.astype(np.float32), rays_d=g_rd.astype(np.float32), ray_o=g[0], ray_d=g[1], near=g[2], far=g[3],
                        mask_at_box=g[4])

    # ---------------- network pieces (torch CPU) ----------------------------------------------
    sd = synthetic.make_state_dict(seed=0, num_train_frame=cfg.num_train_frame)
    net = ref.Network()
    net.load_state_dict(sd)
    net.train()
    renderer = ref.Renderer(net)
    cfg.perturb = 0.

    small = synthetic.make_frame(pose_seed=2, body_seed=1, voxel=0.1, latent_index=3)
    Kc, Rc, Tc = synthetic.make_camera(small, 96, 96, focal=100.0)
    ray_o, ray_d, near, far, mask = ref.dutils.get_rays_within_bounds(96, 96, Kc, Rc, Tc, small['wbounds'])
    batch = synthetic.make_render_batch(small, ray_o, ray_d, near, far)
    n_rays = ray_o.shape[0]

    with torch.no_grad():
        pts_r, z_r = renderer.get_wsampling_points(batch['ray_o'], batch['ray_d'], batch['near'], batch['far'])
        pts_o, z_o = O.sample_points(batch['ray_o'], batch['ray_d'], batch['near'], batch['far'], 64)
        report('get_wsampling_points', 'tpose_renderer.py:14-39', biteq(pts_r.numpy(), pts_o.numpy()) and biteq(z_r.numpy(), z_o.numpy()), 'bit-equal')
        wpts = pts_r.view(1, -1, 3)
        pp_r = ref.blend_utils.world_points_to_pose_points(wpts, batch['R'], batch['Th'])
        pp_o = O.world_to_pose(wpts, batch['R'], batch['Th'])
        report('world_points_to_pose_points', 'blend_utils.py:6-16', biteq(pp_r.numpy(), pp_o.numpy()), 'bit-equal')
        bw_r = ref.blend_utils.pts_sample_blend_weights(pp_r, batch['pbw'], batch['pbounds'])
        bw_o = O.sample_blend_weights(pp_r, batch['pbw'], batch['pbounds'])
        report('pts_sample_blend_weights', 'blend_utils.py:119-149', biteq(bw_r.numpy(), bw_o.numpy()), 'bit-equal')
        sub = pp_r[:, :4096]
        smpl = bw_r[:, :24, :4096]
        idx = batch['latent_index'] + 1
        nb_r = net.calculate_neural_blend_weights(sub, smpl, idx)
        nb_o = O.neural_blend_weights(sd, sub, smpl, idx)
        report('calculate_neural_blend_weights', 'tpose_nerf_network.py:55-77', biteq(nb_r.numpy(), nb_o.numpy()), 'bit-equal')
        tp_r = ref.blend_utils.pose_points_to_tpose_points(sub, nb_r, batch['A'])
        tp_o = O.inverse_lbs(sub, nb_r, batch['A'])
        report('pose_points_to_tpose_points', 'blend_utils.py:41-59', biteq(tp_r.numpy(), tp_o.numpy()), 'bit-equal')
        fw_r = ref.blend_utils.tpose_points_to_pose_points(tp_r, nb_r, batch['A'])
        fw_o = O.forward_lbs(tp_r, nb_r, batch['A'])
        report('tpose_points_to_pose_points', 'blend_utils.py:77-90', biteq(fw_r.numpy(), fw_o.numpy()), 'bit-equal')
        vd = batch['ray_d'][:, :64].repeat(1, 64, 1)
        a_r, c_r = net.tpose_human.calculate_alpha_rgb(tp_r, vd, batch['latent_index'])
        a_o, c_o = O.nerf_alpha_rgb(sd, tp_r, vd, batch['latent_index'])
        report('TPoseHuman.calculate_alpha_rgb', 'tpose_nerf_network.py:252-275', biteq(a_r.numpy(), a_o.numpy()) and biteq(c_r.numpy(), c_o.numpy()), 'bit-equal')
        pe_r = ref.embedder.xyz_embedder(sub)
        report('xyz_embedder', 'embedder.py:5-54', biteq(pe_r.numpy(), O.positional_encoding(sub, 10).numpy()), 'bit-equal')
        raw = torch.rand(n_rays, 64, 4)
        zz = z_r.view(-1, 64)
        r2_r = ref.nerf_net_utils.raw2outputs(raw, zz, False)
        r2_o = O.raw2outputs(raw, zz, False)
        report('raw2outputs', 'nerf_net_utils.py:6-36', all(biteq(a.numpy(), b.numpy()) for a, b in zip(r2_r, r2_o)), 'bit-equal (5 outputs)')
        ga_r = net.get_alpha(wpts[0, :20000], batch)
        ga_o = O.calculate_alpha(sd, wpts[0, :20000], batch, O.OracleCfg())
        report('Network.calculate_alpha', 'tpose_nerf_network.py:105-137', biteq(ga_r.numpy(), ga_o.numpy()), 'bit-equal')

        # full render: reference vs oracle
        out_r = renderer.render(batch)
        out_o = O.render(sd, batch, O.OracleCfg(perturb=0.))
        keys = ['rgb_map', 'acc_map', 'depth_map', 'raw', 'pbw', 'tbw']
        ok = all(biteq(out_r[k].numpy(), out_o[k].numpy()) for k in keys)
        report(f'Renderer.render ({n_rays} rays, {(n_rays + 2047) // 2048} chunks)', 'tpose_renderer.py:159-186', ok,
               'all 6 outputs bit-equal; pbw rows %d' % out_r['pbw'].shape[1])

        # jittered sampling (perturb > 0, net.training) with the reference's RNG stream
        cfg.perturb = 1.
        torch.manual_seed(5)
        out_rj = renderer.render(batch)
        torch.manual_seed(5)
        t_rand = torch.cat([torch.rand(1, min(2048, n_rays - i), 64) for i in range(0, n_rays, 2048)], dim=1)
        out_oj = O.render(sd, batch, O.OracleCfg(perturb=1.), t_rand=t_rand)
        report('Renderer.render with stratified jitter', 'tpose_renderer.py:29-37',
               all(biteq(out_rj[k].numpy(), out_oj[k].numpy()) for k in keys), 'bit-equal given the same t_rand')
        cfg.perturb = 0.

    buf = {k: out_r[k].numpy() for k in keys}
    np.savez_compressed(os.path.join(GOLDEN, 'render_small.npz'),
                        sd_seed=0, sd_digest=sd_digest(sd), voxel=0.1, pose_seed=2, body_seed=1, latent_index=3,
                        K=Kc, R=Rc, T=Tc, H=96, W=96,
                        ray_o=ray_o, ray_d=ray_d, near=near, far=far,
                        **{'frame_' + k: np.asarray(small[k]) for k in synthetic.FRAME_KEYS}, **buf,
                        alpha_grid=ga_r.numpy(), jitter_t_rand=t_rand.numpy(),
                        **{'jitter_' + k: out_rj[k].numpy() for k in ('rgb_map', 'acc_map', 'depth_map')})

    # ---------------- silhouette-culled renderer (tpose_renderer_mmsk) --------------------------
    mb = synthetic.add_silhouettes(batch, small, n_views=4)
    with torch.no_grad():
        mren = ref.MmskRenderer(net)
        ins_r = mren.prepare_inside_pts(pts_r[:, :2048], mb)
        ins_o = O.inside_all_views(pts_r[:, :2048].reshape(1, -1, 3), mb)
        report('Renderer.prepare_inside_pts (4 views)', 'tpose_renderer_mmsk.py:14-57', biteq(ins_r.numpy(), ins_o.numpy()),
               f'bit-equal; {int(ins_r.sum())} of {ins_r.numel()} samples inside every silhouette')
        out_rm = mren.render(mb)
        out_om = O.render_mmsk(sd, mb, O.OracleCfg(perturb=0.), return_debug=True)
    mkeys = ('rgb_map', 'acc_map', 'depth_map')
    report('tpose_renderer_mmsk.Renderer.render', 'tpose_renderer_mmsk.py:99-166',
           all(biteq(out_rm[k].numpy(), out_om[k].numpy()) for k in mkeys) and set(out_rm.keys()) == set(mkeys),
           f'rgb/acc/depth bit-equal; {int(out_om["_debug"]["inside"].sum())} samples survive the culling')
    np.savez_compressed(os.path.join(GOLDEN, 'render_small_mmsk.npz'), n_views=4, msks=mb['msks'][0].numpy(), Ks=mb['Ks'][0].numpy(),
                        RT=mb['RT'][0].numpy(), H=int(mb['H']), W=int(mb['W']), inside=out_om['_debug']['inside'].numpy(),
                        **{k: out_rm[k].numpy() for k in mkeys})

    # ---------------- novel-view camera path (render_utils.gen_path) ------------------------------
    if ref.render_utils is not None:
        from animatable_nerf_b200 import host_geometry
        rig = [m.astype(np.float64) for m in mb['RT'][0].numpy()]
        cfg.render_views = 16
        path_r = np.array(ref.render_utils.gen_path([m.copy() for m in rig]))
        path_o = np.array(host_geometry.circular_camera_path([m.copy() for m in rig], 16))
        report('gen_path (16-view circular sweep)', 'render_utils.py:77-127', path_r.shape == path_o.shape and float(np.abs(path_r - path_o).max()) < 1e-12,
               f'max abs diff {float(np.abs(path_r - path_o).max()):.1e} (float64 host code)')

    # ---------------- training step (tpose_trainer.NetworkWrapper + Trainer.train up to the optimizer) ---------
    if ref.trainer_mod is not None:
        tb, t_rand = synthetic.make_train_batch(small, ray_o, ray_d, near, far, n_rays=512)
        cfg.perturb = 1.
        net.train()
        net.zero_grad()
        cwd = os.getcwd()
        os.chdir(reference_import.REF)          # make_renderer loads cfg.renderer_path relative to the reference root
        try:
            wrapper = ref.trainer_mod.NetworkWrapper(net)
        finally:
            os.chdir(cwd)
        torch.manual_seed(5)
        # the reference draws the jitter with torch.rand on the global CPU generator (tpose_renderer.py:35)
        t_rand = torch.rand(1, tb['ray_o'].shape[1], 64)
        torch.manual_seed(5)
        _, loss_r, stats_r, _ = wrapper(tb)
        loss_r.mean().backward()
        torch.nn.utils.clip_grad_value_(net.parameters(), 40)
        grads_r = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in net.named_parameters()}
        stats_o, grads_o = O.train_step_grads(sd, tb, O.OracleCfg(perturb=1.), t_rand=t_rand)
        ok = all(biteq(grads_r[k].numpy(), grads_o[k].numpy()) for k in grads_r) and \
            all(float(stats_r[k]) == stats_o[k] for k in ('bw_loss', 'img_loss', 'loss'))
        worst = max(maxdiff(grads_r[k], grads_o[k]) for k in grads_r)
        report('tpose_trainer step: loss + 46 parameter gradients', 'tpose_trainer.py:21-73, trainer.py:62-66', ok,
               f'bit-equal (loss {stats_o["loss"]:.6f} = bw {stats_o["bw_loss"]:.3e} + img {stats_o["img_loss"]:.6f}); max grad diff {worst:.1e}')
        np.savez_compressed(os.path.join(GOLDEN, 'train_step_small.npz'), n_rays=512, t_rand=t_rand.numpy(),
                            ray_o=tb['ray_o'][0].numpy(), ray_d=tb['ray_d'][0].numpy(), near=tb['near'][0].numpy(), far=tb['far'][0].numpy(),
                            rgb=tb['rgb'][0].numpy(), mask_at_box=tb['mask_at_box'][0].numpy(),
                            **{'stat_' + k: np.float32(float(stats_r[k])) for k in ('bw_loss', 'img_loss', 'loss')},
                            **{'grad_' + k: grads_r[k].numpy() for k in ('bw_fc.weight', 'bw_fc.bias', 'bw_linears.0.bias', 'bw_latent.weight',
                                                                         'tpose_human.alpha_fc.weight', 'tpose_human.rgb_fc.weight',
                                                                         'tpose_human.pts_linears.0.bias', 'tpose_human.pts_linears.7.bias',
                                                                         'tpose_human.nf_latent.weight', 'tpose_human.view_fc.bias')},
                            **{'gradnorm_' + k: np.float32(float(grads_r[k].norm())) for k in grads_r})
        cfg.perturb = 0.
        net.zero_grad()

    # ---------------- novel-pose field (aninerf_s9p stage 2 shapes) -----------------------------
    sd2 = synthetic.make_state_dict(seed=1, num_train_frame=cfg.num_train_frame, num_eval_frame=8)
    cfg.aninerf_animation = True
    cfg.num_eval_frame = 8
    cfg.test_novel_pose = True
    if 'init_aninerf' in cfg:
        cfg.pop('init_aninerf')
    net2 = ref.Network()
    net2.load_state_dict(sd2)
    net2.train()
    with torch.no_grad():
        out_r2 = ref.Renderer(net2).render(batch)
        out_o2 = O.render(sd2, batch, O.OracleCfg(perturb=0., test_novel_pose=True))
    report('Renderer.render, test_novel_pose (novel_pose_bw field)', 'tpose_nerf_network.py:93-94,278-315',
           all(biteq(out_r2[k].numpy(), out_o2[k].numpy()) for k in keys), 'all 6 outputs bit-equal')
    np.savez_compressed(os.path.join(GOLDEN, 'render_small_novel_pose.npz'), sd_seed=1, sd_digest=sd_digest(sd2), num_eval_frame=8,
                        **{k: out_r2[k].numpy() for k in ('rgb_map', 'acc_map', 'depth_map')})
    # ---------------- stage 2: novel-pose blend-weight training (aninerf_animation_trainer) -----------------
    if ref.anim_trainer_mod is not None:
        import unittest.mock as mock
        n_pts = 4096
        gen = torch.Generator().manual_seed(6)
        draws = [torch.rand(1, n_pts, generator=gen) for _ in range(6)]
        wb, tbd = batch['wbounds'], batch['tbounds']
        wpts = (wb[:, 1] - wb[:, 0])[:, None] * torch.stack(draws[0:3], dim=2) + wb[:, 0][:, None]
        tpts_s = (tbd[:, 1] - tbd[:, 0])[:, None] * torch.stack(draws[3:6], dim=2) + tbd[:, 0][:, None]
        it = iter(draws)
        cwd = os.getcwd()
        os.chdir(reference_import.REF)
        try:
            wrapper2 = ref.anim_trainer_mod.NetworkWrapper(net2)
        finally:
            os.chdir(cwd)
        net2.zero_grad()
        # the reference draws 1024*64 points with torch.rand; feed it our (smaller) seeded draws instead
        with mock.patch.object(ref.anim_trainer_mod.torch, 'rand', side_effect=lambda *a, **k: next(it)):
            _, loss2, stats2, _ = wrapper2(batch)
        loss2.mean().backward()
        torch.nn.utils.clip_grad_value_(net2.parameters(), 40)
        g_r = {k: p.grad.clone() for k, p in net2.named_parameters() if p.grad is not None}
        st_o, g_o = O.animation_train_step_grads(sd2, batch, wpts, tpts_s, O.OracleCfg())
        ok = set(g_r) == set(g_o) and all(biteq(g_r[k].numpy(), g_o[k].numpy()) for k in g_r) and \
            all(float(stats2[k]) == st_o[k] for k in ('bw_loss0', 'bw_loss1', 'loss'))
        report('aninerf_animation_trainer step: 2 losses + novel_pose_bw gradients', 'aninerf_animation_trainer.py:33-119', ok,
               f'bit-equal ({len(g_r)} trainable tensors; loss {st_o["loss"]:.6e} = {st_o["bw_loss0"]:.3e} + {st_o["bw_loss1"]:.3e})')
        np.savez_compressed(os.path.join(GOLDEN, 'animation_train_step_small.npz'), wpts=wpts[0].numpy(), tpts=tpts_s[0].numpy(),
                            **{'stat_' + k: np.float32(float(stats2[k])) for k in ('bw_loss0', 'bw_loss1', 'loss')},
                            **{'grad_' + k: g_r[k].numpy() for k in ('novel_pose_bw.bw_fc.weight', 'novel_pose_bw.bw_fc.bias',
                                                                     'novel_pose_bw.bw_linears.0.bias', 'novel_pose_bw.bw_latent.weight')},
                            **{'gradnorm_' + k: np.float32(float(g_r[k].norm())) for k in g_r})
        for prm in net2.parameters():
            prm.requires_grad = True
    cfg.test_novel_pose = False
    cfg.aninerf_animation = False

    # ---------------- report ---------------------------------------------------------------------
    with open(os.path.join(ROOT, 'oracle', 'VALIDATION.md'), 'w') as f:
        f.write('# Oracle pinned against the unmodified reference\n\n')
        f.write(f'Generated by `python -m oracle.validate_against_reference` in the build container '
                f'(torch {torch.__version__}, numpy {np.__version__}, CPU, {os.cpu_count()} threads).\n'
                'Reference imported unmodified from /root/reference with the module stubs of SURVEY.md section 8c.\n'
                'The reference ships no tests/golden vectors of its own (SURVEY.md section 4); the fixtures under\n'
                '`tests/golden/` are REFERENCE outputs produced by this script.\n\n')
        f.write('| oracle function vs reference | reference file:line | result | detail |\n|---|---|---|---|\n')
        for r in rows:
            f.write('| ' + ' | '.join(r) + ' |\n')
    bad = [r for r in rows if r[2] != 'PASS']
    print(f'{len(rows) - len(bad)}/{len(rows)} checks passed')
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
