"""Numpy restatement of the reference's adaptive-refinement decision (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

EvaluateBlock (reference radiation_adaptive.cpp:163-312): five criteria on the Stokes-I values of one block at
adaptive_frequency_num -- value, absolute / relative gradient (one-sided at the block edges), absolute / relative
Laplacian (interior pixels only) -- each "refine if the fraction of finite pixels exceeding the cut is larger than the
given fraction", evaluated in order, a negative fraction disabling a criterion.  Forced regions
(radiation_adaptive.cpp:54-66) refine a block whose centre lies in the rectangle while the level is below the region's.
child_locs lists the children the way AugmentCamera orders them (camera.cpp:445-459).
"""
import numpy as np


def _exceeds(q, cut, frac):
    finite = np.isfinite(q)
    examined = int(finite.sum())
    exceeded = int((q[finite] > cut).sum())
    with np.errstate(all='ignore'):
        return np.float64(exceeded) / np.float64(examined) > frac     # 0/0 = nan compares false, as in C


def evaluate_block(I, kv):
    """I: (bs, bs) intensities of one block, I[i, j] with i the row; kv: the input file's adaptive_* keys."""
    I = np.asarray(I, np.float64)
    get = lambda k: float(kv[k])
    with np.errstate(all='ignore'):
        if get('adaptive_val_frac') >= 0.0 and _exceeds(np.abs(I), get('adaptive_val_cut'), get('adaptive_val_frac')):
            return True
        if get('adaptive_abs_grad_frac') >= 0.0:
            qx = np.empty_like(I)
            qx[:, 0], qx[:, -1] = I[:, 1] - I[:, 0], I[:, -1] - I[:, -2]
            qx[:, 1:-1] = 0.5 * (I[:, 2:] - I[:, :-2])
            qy = np.empty_like(I)
            qy[0], qy[-1] = I[1] - I[0], I[-1] - I[-2]
            qy[1:-1] = 0.5 * (I[2:] - I[:-2])
            if _exceeds(np.hypot(qx, qy), get('adaptive_abs_grad_cut'), get('adaptive_abs_grad_frac')):
                return True
        if get('adaptive_rel_grad_frac') >= 0.0:
            qx = np.empty_like(I)
            qx[:, 0] = 2.0 * (I[:, 1] - I[:, 0]) / (I[:, 0] + I[:, 1])
            qx[:, -1] = 2.0 * (I[:, -1] - I[:, -2]) / (I[:, -2] + I[:, -1])
            qx[:, 1:-1] = 2.0 * (I[:, 2:] - I[:, :-2]) / (I[:, :-2] + 2.0 * I[:, 1:-1] + I[:, 2:])
            qy = np.empty_like(I)
            qy[0] = 2.0 * (I[1] - I[0]) / (I[0] + I[1])
            qy[-1] = 2.0 * (I[-1] - I[-2]) / (I[-2] + I[-1])
            qy[1:-1] = 2.0 * (I[2:] - I[:-2]) / (I[:-2] + 2.0 * I[1:-1] + I[2:])
            if _exceeds(np.hypot(qx, qy), get('adaptive_rel_grad_cut'), get('adaptive_rel_grad_frac')):
                return True
        c = I[1:-1, 1:-1]
        lx, ly = I[1:-1, :-2] - 2.0 * c + I[1:-1, 2:], I[:-2, 1:-1] - 2.0 * c + I[2:, 1:-1]
        if get('adaptive_abs_lapl_frac') >= 0.0 and _exceeds(np.abs(lx + ly), get('adaptive_abs_lapl_cut'), get('adaptive_abs_lapl_frac')):
            return True
        if get('adaptive_rel_lapl_frac') >= 0.0:
            qx = 4.0 * lx / (I[1:-1, :-2] + 2.0 * c + I[1:-1, 2:])
            qy = 4.0 * ly / (I[:-2, 1:-1] + 2.0 * c + I[2:, 1:-1])
            if _exceeds(np.abs(qx + qy), get('adaptive_rel_lapl_cut'), get('adaptive_rel_lapl_frac')):
                return True
    return False


def refinement_flags(image, locs, level, kv):
    """image: (B, bs, bs) block images of one level; locs: (B, 2) block (v, u).  Forced regions first, then criteria."""
    bs = int(kv['adaptive_block_size'])
    n_blocks = int(kv['camera_resolution']) // bs * 2 ** level
    width = float(kv['camera_width'])
    flags = np.zeros(len(locs), np.uint8)
    for b, (v, u) in enumerate(locs):
        forced = False
        for n in range(1, int(kv.get('adaptive_num_regions', 0)) + 1):
            if level < int(kv['adaptive_region_%d_level' % n]):
                x, y = ((u + 0.5) / n_blocks - 0.5) * width, ((v + 0.5) / n_blocks - 0.5) * width
                if (float(kv['adaptive_region_%d_x_min' % n]) < x < float(kv['adaptive_region_%d_x_max' % n])
                        and float(kv['adaptive_region_%d_y_min' % n]) < y < float(kv['adaptive_region_%d_y_max' % n])):
                    forced = True
        flags[b] = 1 if forced or evaluate_block(image[b], kv) else 0
    return flags


def child_locs(locs, flags):
    out = []
    for (v, u), f in zip(locs, flags):
        if f:
            out += [(2 * v, 2 * u), (2 * v, 2 * u + 1), (2 * v + 1, 2 * u), (2 * v + 1, 2 * u + 1)]
    return np.array(out, np.int32).reshape(-1, 2)
