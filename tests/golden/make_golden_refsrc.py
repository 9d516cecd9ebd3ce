"""Golden vectors from the reference's OWN source, executed in this image.

    python tests/golden/make_golden_refsrc.py            (needs /root/reference; writes tests/golden/refsrc/*.npz)

The reference (pure Python on JAX) cannot be imported here as it stands: jax / jaxlib are not installable.  This script puts
``tests/_jaxshim`` - a float64 torch-backed stand-in for the ~25 JAX entry points the objective path touches - in front of
``/root/reference/src`` on ``sys.path`` and then imports and calls the reference's unmodified modules:

    eincm.losses.loss_func / handover_loss_func / compute_loss_objectives       (src/eincm/losses.py:49-276)
      -> utils.theta_utils.scale_theta_to_sensor_size, eincm.event_warpers.per_pix_warp, utils.event_utils.events_to_pdf_frame,
         utils.img_utils.normalize_to_unit_range / sobel_scharr_optimized_image_grads, eincm.objectives.*, eincm.regularizers.*

on the INPUTS of the committed fixtures tests/golden/*.npz, forward and (``jax.value_and_grad`` of the stand-in = torch autograd) reverse.
What these vectors pin is the reference's composition of the primitives as executed - not the primitives, which the stand-in restates
from JAX's published behaviour (tests/_jaxshim/jax/__init__.py lists them).  No reference source is copied: it is imported where it lies.

With a real JAX at hand, ``EINCM_REAL_JAX=1 python tests/golden/make_golden_refsrc.py`` (and ..._fullsize.py) runs the same calls on
it and rewrites the vectors; every test that reads them then checks the oracle and the CUDA path against JAX itself, which is what
closes "parity unpinned" (DESIGN.md section 2).
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REF_SRC = os.environ.get('EINCM_REFERENCE_SRC', '/root/reference/src')
OUT = os.path.join(HERE, 'refsrc')


def import_reference():
    """the reference's eincm.losses on top of the stand-in; raises ImportError when the reference tree is absent"""
    if not os.path.isdir(REF_SRC):
        raise ImportError(f'{REF_SRC} not present')
    real = os.environ.get('EINCM_REAL_JAX') == '1'      # a machine with jax / jaxlib (/ jaxopt): the same script, no stand-in
    for p in (REF_SRC,) + (() if real else (os.path.join(TESTS, '_jaxshim'),)):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import jax
    if real:
        jax.config.update('jax_enable_x64', True)       # src/experiments/e00/configs/jax_config/default.yaml:2
    assert real != ('TEST INFRASTRUCTURE ONLY' in (jax.__doc__ or '')), 'stand-in / real jax mix-up'
    import eincm.losses as L
    assert os.path.realpath(L.__file__).startswith(os.path.realpath(REF_SRC))
    return jax, L


def run_case(jax, L, g):
    """g: dict of a tests/golden/*.npz fixture -> dict of reference outputs"""
    hp = {str(n): float(v) for n, v in zip(g['hp_names'], g['hp_values'])}
    H, W = g['edges'].shape[1:]
    import jax.numpy as jnp
    xs, ys, ts, edges, edge_ts = (jnp.array(g[k]) for k in ('xs', 'ys', 'ts', 'edges', 'edge_ts'))
    kw = dict(alpha=hp['alpha'], beta=hp['beta'], gamma=hp['gamma'], delta=hp['delta'], cur_pyr_lvl=int(hp['cur_pyr_lvl']), n_pyr_lvls=5,
              sensor_size=(int(H), int(W)), scale_to_sensor_size_method='bilinear')

    def f(theta):
        return L.loss_func(theta, xs, ys, ts, edges, edge_ts, **kw)
    (loss, aux), grad = jax.value_and_grad(f, has_aux=True)(jnp.array(g['theta']))
    obj = L.compute_loss_objectives(aux['scaled_theta'], xs, ys, ts, edges, edge_ts, (int(H), int(W)))
    a = float(g['alpha_handover'])

    def ho(alpha_handover):
        return L.handover_loss_func(alpha_handover, jnp.array(g['prev_theta']), jnp.array(g['theta']), xs, ys, ts, edges, edge_ts, **kw)
    ho_loss, ho_dalpha = jax.value_and_grad(ho)(jnp.array(a))
    out = dict(loss=float(loss), grad=np.asarray(grad), scaled_theta=np.asarray(aux['scaled_theta']),
               mean_rel_corr=float(aux['mean_rel_corr']), mean_rel_contrast=float(aux['mean_rel_contrast']),
               mean_rel_iwe_divergence=float(aux['mean_rel_iwe_divergence']), theta_total_variation_used=float(aux['theta_total_variation']),
               handover_loss=float(ho_loss), handover_dalpha=float(ho_dalpha))
    for k, v in obj.items():
        out['obj_' + k] = np.asarray(v, dtype=np.float64)
    # the images themselves (losses.py:54, 61: what compute_loss_objectives builds and reduces)
    out['zero_iwe'] = np.asarray(L.events_to_pdf_frame(xs, ys, (int(H), int(W))))
    out['iwes'] = np.asarray(L.vmapped_events_to_pdf_frame(obj['warped_xs'], obj['warped_ys'], (int(H), int(W))))
    if 'gt_flow' in g:
        # src/evaluations/theta_eval.py:14-97 (and flow_eval.py:14-75 through it) on the fixture's synthetic ground truth
        import evaluations.theta_eval as TE
        _, _, ev, _ = TE.evaluate_theta_array(aux['scaled_theta'], xs, ys, ts, edges, edge_ts, jnp.array(g['gt_flow']), hp['alpha'], hp['beta'],
                                              hp['gamma'], hp['delta'], (int(H), int(W)), err_eval_event_mask=jnp.array(g['err_mask']))
        for k, v in ev.items():
            out['eval_' + k] = np.asarray(v, dtype=np.float64)
    return out


def main():
    jax, L = import_reference()
    os.makedirs(OUT, exist_ok=True)
    for path in sorted(glob.glob(os.path.join(HERE, '*.npz'))):
        z = np.load(path)
        g = {k: z[k] for k in z.files}
        out = run_case(jax, L, g)
        name = os.path.basename(path)
        # the per-event warped coordinates are the bulky part: keep them as float64 (bit-exact index stream), drop nothing else
        np.savez_compressed(os.path.join(OUT, name), **out)
        print(f'{name}: loss {out["loss"]:.12g} (oracle fixture {float(g["loss"]):.12g})  |grad|inf {np.abs(out["grad"]).max():.6g}'
              f'  grad diff {np.abs(out["grad"] - g["grad"]).max():.3g}')


if __name__ == '__main__':
    main()
