"""Parameter presets shared by the tests, bench.py and the bindings.

Mirrors what av1/encoder/speed_features.c derives from --cpu-used / frame size
for the temporal-filter path: sub-pel method, iterations per
step, mesh-prune level, mesh pattern row, skip-row SAD.
"""

MESH_ROWS = {  # speed_features.c:25-35
    0: [(64, 8), (28, 4), (15, 1), (7, 1)],
    1: [(64, 8), (28, 4), (15, 1), (7, 1)],
    2: [(64, 8), (14, 2), (7, 1), (7, 1)],
    3: [(64, 16), (24, 8), (12, 4), (7, 1)],
    4: [(64, 16), (24, 8), (12, 4), (7, 1)],
    5: [(64, 16), (24, 8), (12, 4), (7, 1)],
}


def tf_params(width, height, num_frames, filter_frame_idx=None, bit_depth=8, use_hbd=None, ss_x=1, ss_y=1,
              monochrome=0, speed=4, q_factor=32, filter_strength=5, noise_levels=(2.0, 1.0, 1.0),
              allow_hp=0, force_integer_mv=0, compute_frame_diff=1, border=None, **over):
    if use_hbd is None:
        use_hbd = 1 if bit_depth > 8 else 0
    if filter_frame_idx is None:
        filter_frame_idx = num_frames // 2
    if border is None:
        # encoder_utils.h:1103-1115: sb_size + 32; SB64 at speed>=1 for <=480p (encoder_utils.c:825)
        border = (64 if (min(width, height) <= 480 and speed >= 1) else 128) + 32
    is720 = min(width, height) >= 720
    p = dict(
        width=width, height=height, ss_x=ss_x, ss_y=ss_y, monochrome=monochrome, bit_depth=bit_depth,
        use_hbd=use_hbd, border=border, num_frames=num_frames, filter_frame_idx=filter_frame_idx,
        noise_levels=tuple(noise_levels), q_factor=q_factor, filter_strength=filter_strength,
        force_integer_mv=force_integer_mv, allow_hp=allow_hp,
        # speed_features.c:1905 (TREE), :1070 (speed>=3 PRUNED), :1134 (speed>=4 PRUNED_MORE)
        subpel_method=0 if speed <= 2 else (1 if speed == 3 else 2),
        # speed_features.c:1904 (2), :1007 (1 from speed 2 below 720p) -- see SURVEY App.A 12
        subpel_iters_per_step=2 if speed <= 1 else 1,
        # speed_features.c:1899 (off), :1073 (LVL_1 speed 3), :1167 (LVL_2 speed>=4)
        prune_mesh_level=0 if speed <= 2 else (1 if speed == 3 else 2),
        mesh=MESH_ROWS[min(speed, 5)],
        # speed_features.c:619-623
        use_downsampled_sad=1 if is720 else 0,
        compute_frame_diff=compute_frame_diff,
    )
    p.update(over)
    return p
