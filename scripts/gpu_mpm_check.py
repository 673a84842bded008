"""GPU bring-up of the MPM solver: parity against the C oracle (fp32 and fp64), then timing of
BASELINE config 3 (100k particles, 64^3 grid, 1000 substeps). Run under gpurun."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mpm_ref as R
from pixie_b200.mpm_solver_warp import MPM_Simulator_WARP

def log(*a): print(*a, flush=True)

def build_pair(n, ng, materials, seed, with_bcs=True, prec="f32", parallel=0):
    sc = R.synthetic_scene(n, ng, seed=seed, materials=materials)
    dev = "cuda:0"
    s = MPM_Simulator_WARP(10)
    s.load_initial_data_from_torch(torch.from_numpy(sc["x"]).to(dev), torch.from_numpy(sc["vol"]).to(dev), None, n_grid=ng, grid_lim=2.0)
    s.set_parameters_dict({"material": "jelly", "g": [0.0, 0.0, -9.8], "density": 1000.0, "E": 1e5, "nu": 0.3,
                           "yield_stress": 2e3, "grid_v_damping_scale": 0.9999, "rpic_damping": 0.0, "friction_angle": 30.0,
                           "hardening": 1, "xi": 0.1, "softening": 0.1, "plastic_viscosity": 10.0, "bulk_modulus": 1e5})
    # per-particle fields as the material-field transfer would set them
    s.mpm_model.E = torch.from_numpy(sc["E"]).to(dev)
    s.mpm_model.nu = torch.from_numpy(sc["nu"]).to(dev)
    s.mpm_state.particle_material = torch.from_numpy(sc["material"]).to(dev)
    s.reset_densities_and_update_masses(torch.from_numpy(sc["density"]).to(dev))
    s.import_particle_v_from_torch(torch.from_numpy(sc["v"]).to(dev))
    s.finalize_mu_lam()

    o = R.MpmRef(n, ng, 2.0, prec)
    o.set("X", sc["x"]); o.set("V", sc["v"]); o.set("VOL", sc["vol"]); o.set("DENSITY", sc["density"])
    o.set("E", sc["E"]); o.set("NU", sc["nu"]); o.set("MATERIAL", sc["material"])
    o.set("YIELD", np.full(n, 2e3)); o.set("BULK", np.full(n, 1e5))
    o.compute_mass(); o.compute_mu_lam()
    o.set_params(g=(0, 0, -9.8), grid_v_damping_scale=0.9999, rpic_damping=0.0, alpha=R.friction_alpha(30.0),
                 hardening=1, xi=0.1, softening=0.1, plastic_viscosity=10.0, parallel_p2g=parallel)
    if with_bcs:
        s.add_bounding_box(); o.add_bc(R.BC_BBOX)
        s.set_velocity_on_cuboid(point=[1.0, 1.0, 0.62], size=[0.5, 0.5, 0.04], velocity=[0, 0, 0]); o.add_bc(R.BC_CUBOID, point=[1.0, 1.0, 0.62], size=[0.5, 0.5, 0.04], velocity=[0, 0, 0], end_time=999.0)
        s.set_velocity_on_cuboid(point=[0.7, 1.0, 1.3], size=[0.05, 0.2, 0.05], velocity=[0.5, 0, 0], start_time=0.0, end_time=0.01, reset=1)
        o.add_bc(R.BC_CUBOID, point=[0.7, 1.0, 1.3], size=[0.05, 0.2, 0.05], velocity=[0.5, 0, 0], start_time=0.0, end_time=0.01, reset=1)
        s.add_surface_collider(point=[1.0, 1.0, 0.1], normal=[0, 0, 1], surface="sticky", friction=0.0, start_time=0.0, end_time=1e3)
        o.add_bc(R.BC_SURFACE, point=[1.0, 1.0, 0.1], normal=[0, 0, 1], surface_type=0, end_time=1e3)
        mask = ((np.abs(sc["x"] - np.array([1.0, 1.0, 1.2], dtype=np.float32)) < np.array([0.2, 0.2, 0.1], dtype=np.float32)).all(1)).astype(np.int32)
        s.add_impulse_on_particles(force=[0.05, 0.0, -0.02], dt=1e-4, point=[1.0, 1.0, 1.2], size=[0.2, 0.2, 0.1], num_dt=20, start_time=0.0)
        o.add_bc(R.BC_IMPULSE, point=[1.0, 1.0, 1.2], size=[0.2, 0.2, 0.1], velocity=[0.05, 0.0, -0.02], start_time=0.0, end_time=0.0 + 1e-4 * 20, mask=mask)
        gm = s._masks[-1].cpu().numpy()
        assert (gm == mask).all(), "selection mask mismatch"
        mask2 = ((np.abs(sc["x"] - np.array([1.3, 1.3, 0.9], dtype=np.float32)) < np.array([0.1, 0.1, 0.1], dtype=np.float32)).all(1)).astype(np.int32)
        s.enforce_particle_velocity_translation(point=[1.3, 1.3, 0.9], size=[0.1, 0.1, 0.1], velocity=[0, 0.2, 0], start_time=0.001, end_time=0.004)
        o.add_bc(R.BC_VTRANS, point=[1.3, 1.3, 0.9], size=[0.1, 0.1, 0.1], velocity=[0, 0.2, 0], start_time=0.001, end_time=0.004, mask=mask2)
    return s, o, sc

def compare(s, o, tag):
    torch.cuda.synchronize()
    out = {}
    for name, fid in (("x", "X"), ("v", "V"), ("F", "F"), ("F_trial", "F_TRIAL"), ("C", "C"), ("stress", "STRESS")):
        a = s._t[fid].detach().cpu().numpy().astype(np.float64).reshape(s.n_particles, -1)
        b = o.get(fid).reshape(s.n_particles, -1)
        scale = max(1e-30, np.abs(b).max())
        out[name] = (np.abs(a - b).max(), scale)
    log(f"   {tag}: " + "  ".join(f"{k}: {e:.2e} (max {sc:.2e})" for k, (e, sc) in out.items()))
    return out

def main():
    for mats, label in (((0,), "jelly"), ((0, 1, 2, 3, 5, 6, 4), "all materials"), ((2,), "sand"), ((1,), "metal"), ((5,), "snow")):
        s, o, sc = build_pair(5000, 32, mats, seed=3)
        o64 = build_pair(5000, 32, mats, seed=3, prec="f64")[1]
        log(f"--- {label}: 5000 particles, 32^3 grid, BCs on")
        for chunk in (1, 9, 40, 150):
            s.p2g2p_n(chunk, 1e-4) if chunk > 1 else s.p2g2p(0, 1e-4)
            o.step(chunk, 1e-4); o64.step(chunk, 1e-4)
            compare(s, o, f"after +{chunk:3d} steps vs f32 oracle")
        compare(s, o64, "after 200 steps   vs f64 oracle")
        x32, x64 = o.get("X"), o64.get("X")
        log(f"   oracle f32 vs f64 noise floor on x: {np.abs(x32 - x64).max():.2e};  time gpu={s.time:.6f} oracle={o.time:.6f}")
        cov = s.export_particle_cov_to_torch(); R9 = s.export_particle_R_to_torch()
        log(f"   export cov/R finite: {torch.isfinite(cov).all().item()} {torch.isfinite(R9).all().item()}")

    # ---- config 3 timing + 1000-step drift
    n, ng, steps = 100_000, 64, 1000
    s, o64, sc = build_pair(n, ng, (0,), seed=0, with_bcs=True, prec="f64", parallel=1)
    for _ in range(3): s.p2g2p_n(25, 1e-4)
    torch.cuda.synchronize()
    # re-create for a clean 1000-step rollout
    s, o64, sc = build_pair(n, ng, (0,), seed=0, with_bcs=True, prec="f64", parallel=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.p2g2p_n(25, 1e-4); torch.cuda.synchronize()      # graph instantiation outside the timed region
    s, o64, sc = build_pair(n, ng, (0,), seed=0, with_bcs=True, prec="f64", parallel=1)
    s.p2g2p_n(25, 1e-4)
    e0.record(); s.p2g2p_n(steps - 25, 1e-4); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    log(f"TIME config3: {steps-25} substeps in {ms:.2f} ms -> {ms/(steps-25)*1e3:.2f} us/substep, {n*(steps-25)/ms*1e3:.3e} particle-steps/s")
    t = time.time(); o64.step(steps, 1e-4); tcpu = time.time() - t
    log(f"CPU oracle f64 ({o64.num_threads()} threads): {steps} substeps in {tcpu:.1f}s -> {n*steps/tcpu:.3e} particle-steps/s")
    x = s._t["X"].cpu().numpy().astype(np.float64)
    log(f"DRIFT after {steps} substeps vs f64 oracle: max |dx| = {np.abs(x - o64.get('X')).max():.3e}  (x range {x.min():.3f}..{x.max():.3f})")

if __name__ == "__main__":
    main()
