"""Host-side scene construction (numpy): block lattices, dummy-particle wall boxes, material tables.

Mirror of the scene builders of the reference's eng/particle_func.py (function names and argument meaning kept):
  calc_cube_particle_num pf:253-259, add_cube pf:176-208, add_cube_boundary pf:210-211, add_particles pf:159-174,
  set_material / get_material pf:218-244, chk_block_in_domain pf:246-251, count_boundary / add_boundary pf:302-313,
  calc_dummy_boundary pf:316-343.
The particle SET (positions, creation order) must be bit-identical to the reference's: it is the precondition of
every parity test.  calc_rep_boundary pf:346-374, load_body pf:265-294 (needs trimesh, like the reference).
"""
import numpy as np

MAT_FLUID, MAT_SOIL, MAT_RIGID, MAT_DUMMY, MAT_REP = 1, 2, 11, -1, -2


def calc_cube_particle_num(translation, size, dim=3, offset=None):
    """Per-axis lattice coordinates ``arange(t + off/2, t + size + 1e-5, off)`` and their product count."""
    step = 0.1 if offset is None else offset
    axes = []
    for a in range(dim):
        s = step if size[a] >= 0 else -step
        axes.append(np.arange(translation[a] + s / 2.0, translation[a] + size[a] + 1e-5, s))
    count = 1
    for ax in axes:
        count *= len(ax)
    return count, axes


def count_cube_num(pos_bld, pos_fru, dim, offset):
    return calc_cube_particle_num(pos_bld, np.asarray(pos_fru) - np.asarray(pos_bld), dim, offset=offset)[0]


def cube_positions(lower_corner, cube_size, dim, offset):
    """(n, 3) float64 lattice, x slowest / last axis fastest (meshgrid 'ij' order), z = 0 in 2D."""
    n, axes = calc_cube_particle_num(lower_corner, cube_size, dim, offset=offset)
    if dim == 2:
        axes = axes + [np.array([0.0])]
    grid = np.meshgrid(*axes, sparse=False, indexing="ij")
    pos = np.stack([g.reshape(-1) for g in grid], axis=1).astype(np.float64)
    assert pos.shape == (n, 3)
    return pos


def add_particles(ps, object_id, new_particles_num, new_particles_positions, new_particles_velocity,
                  new_particle_density, new_particle_pressure, new_particles_material_id, new_particles_material_type,
                  new_particles_is_dynamic, new_particles_color):
    ps._add_particles(object_id, new_particles_num, new_particles_positions, new_particles_velocity,
                      new_particle_density, new_particle_pressure, new_particles_material_id,
                      new_particles_material_type, new_particles_is_dynamic, new_particles_color)


def add_cube(ps, object_id, lower_corner, cube_size, mat_id=0, mat_type=1, is_dynamic=True, color=(0, 0, 0),
             density=None, pressure=None, velocity=None, offset=None):
    step = ps.particle_diameter if offset is None else offset
    pos = cube_positions(lower_corner, cube_size, ps.dim, step)
    n = len(pos)
    vel = np.zeros_like(pos) if velocity is None else np.tile(np.asarray(velocity, dtype=np.float64), (n, 1))
    add_particles(ps, object_id, n, pos, vel,
                  np.full(n, 0.0 if density is None else density, dtype=np.float64),
                  np.full(n, 0.0 if pressure is None else pressure, dtype=np.float64),
                  np.full(n, mat_id, dtype=np.int32), np.full(n, mat_type, dtype=np.int32),
                  np.full(n, int(is_dynamic), dtype=np.int32),
                  np.tile(np.asarray(color, dtype=np.float32), (n, 1)))


def add_cube_boundary(ps, pos_bld, pos_fru, type, obj_id, offset=None, color=(0, 0, 0)):
    add_cube(ps=ps, lower_corner=pos_bld, cube_size=np.asarray(pos_fru) - np.asarray(pos_bld), mat_type=type,
             color=color, object_id=obj_id, offset=offset)


def set_material(ps):
    """Material tables split by kind; ``mat_index[matId] = [kind, index inside that kind's list]``."""
    tables = {ps.mat_fluid_type: [], ps.mat_soil_type: [], ps.mat_rigid_type: []}
    mat_index = []
    for mat in ps.cfg.get_materials():
        kind = mat["matType"]
        if kind in tables:
            mat_index.append([kind, len(tables[kind])])
            tables[kind].append(mat)
    return mat_index, tables[ps.mat_fluid_type], tables[ps.mat_soil_type], tables[ps.mat_rigid_type]


def get_material(ps, i_mat_index):
    kind, k = ps.mat_index[i_mat_index]
    return {ps.mat_fluid_type: ps.mat_fluid, ps.mat_soil_type: ps.mat_soil, ps.mat_rigid_type: ps.mat_rigid}[kind][k]


def chk_block_in_domain(domain_start, domain_end, block_translation, block_size, dim):
    inside = all(block_translation[a] - domain_start[a] >= 0.0 and
                 block_translation[a] + block_size[a] - domain_end[a] <= 0.0 for a in range(dim))
    assert inside, "Block is not in domain!"


def count_boundary(boundary, dim, offset):
    return sum(count_cube_num(lo, hi, dim, offset) for lo, hi in boundary)


def add_boundary(ps, boundary, type, offset=None, color=(255, 255, 255)):
    rgb = np.asarray(color) / 255
    for lo, hi in boundary:
        add_cube_boundary(ps, lo, hi, type, type, offset, rgb)


def calc_dummy_boundary(dim, domain_start, domain_end, vdomain_start, vdomain_end):
    """Wall boxes [lower corner, upper corner], no lid: 2D left / bottom / right; 3D b, r, f, l, d (pf:316-343)."""
    ds, de, vs, ve = (np.asarray(a, dtype=np.float64) for a in (domain_start, domain_end, vdomain_start, vdomain_end))
    if dim == 3:
        return [[np.array([vs[0], ds[1], vs[2]]), np.array([ds[0], de[1], de[2]])],
                [np.array([vs[0], ds[1], de[2]]), np.array([de[0], de[1], ve[2]])],
                [np.array([de[0], ds[1], ds[2]]), np.array([ve[0], de[1], ve[2]])],
                [np.array([ds[0], ds[1], vs[2]]), np.array([ve[0], de[1], ds[2]])],
                [vs.copy(), np.array([ve[0], ds[1], ve[2]])]]
    return [[np.array([vs[0], ds[1], ds[2]]), np.array([ds[0], de[1], de[2]])],
            [np.array([vs[0], vs[1], ds[2]]), np.array([ve[0], ds[1], de[2]])],
            [np.array([de[0], ds[1], ds[2]]), np.array([ve[0], de[1], de[2]])]]


def calc_rep_boundary(dim, domain_start, domain_end, pt_radius):
    """Boxes of the repulsive particles (pf:346-374): ONE layer on every domain face, no lid, half a radius thick on
    either side of the face; they are filled with a spacing of one particle radius (ps:147-148)."""
    ds, de = (np.asarray(a, dtype=np.float64) for a in (domain_start, domain_end))
    t = pt_radius / 2
    if dim == 3:
        return [[np.array([ds[0] - t, ds[1] + t, ds[2] - t]), np.array([ds[0] + t, de[1] - t, de[2] - t])],
                [np.array([ds[0] - t, ds[1] + t, de[2] - t]), np.array([de[0] - t, de[1] - t, de[2] + t])],
                [np.array([de[0] - t, ds[1] + t, ds[2] + t]), np.array([de[0] + t, de[1] - t, de[2] + t])],
                [np.array([ds[0] + t, ds[1] + t, ds[2] - t]), np.array([de[0] + t, de[1] - t, ds[2] + t])],
                [ds - t, np.array([de[0], ds[1], de[2]]) + t]]
    return [[np.array([ds[0] - t, ds[1] + t, ds[2]]), np.array([ds[0] + t, de[1] - t, de[2]])],
            [np.array([ds[0] - t, ds[1] - t, ds[2]]), np.array([de[0] + t, ds[1] + t, de[2]])],
            [np.array([de[0] - t, ds[1] + t, ds[2]]), np.array([de[0] + t, de[1] - t, de[2]])]]


def load_body(body, vox_len):
    """Voxelised points of a mesh body (pf:265-294).  Like the reference this is ``trimesh`` from start to end -- load,
    scale, rotation about ``rotationAxis`` through the reference's pivot, filled voxelisation at one particle diameter --
    so it needs that package at run time (it is not part of this image; no scene of the reference ships a mesh)."""
    try:
        import trimesh as tm
    except ImportError as e:                             # the reference fails at the same line (pf:266)
        raise ImportError("mesh Bodies need the 'trimesh' package, as in the reference (eng/particle_func.py:266)") from e
    mesh = tm.load(body["geometryFile"])
    mesh.apply_scale(body["scale"])
    offset = np.array(body["translation"])
    angle = body["rotationAngle"] / 180 * np.pi
    mesh.apply_transform(tm.transformations.rotation_matrix(angle, body["rotationAxis"], mesh.vertices.mean(axis=-1)))
    if body["isDynamic"]:                                # kept for exporters (pf:279-285)
        backup = mesh.copy()
        backup.vertices += offset
        body["mesh"], body["restPosition"], body["restCenterOfMass"] = backup, backup.vertices, backup.vertices.mean(axis=-1)
    return mesh.voxelized(pitch=vox_len).fill().points + offset
