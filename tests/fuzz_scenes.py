"""Random mixed scenes for parity fuzzing: every primitive type, Translate / RotateY / Translate(RotateY) instances,
constant media over spheres and instanced cubes, BVHs (with duplicated and single primitives), nested lists,
all five materials and four textures.  The same call sequence drives any backend (product, oracle, harness)."""
import numpy as np


def build_random_scene(s, seed):
    rs = np.random.RandomState(seed)
    f = lambda lo, hi: float(rs.uniform(lo, hi))

    img = rs.randint(0, 256, size=(8, 16, 3)).astype(np.uint8)
    tex = [s.texture_solid(f(0, 1), f(0, 1), f(0, 1)) for _ in range(3)]
    tex.append(s.texture_checker(f(0.5, 3.0), tex[0], tex[1]))
    tex.append(s.texture_marble(f(0.1, 2.0), int(rs.randint(0, 1 << 30))))
    tex.append(s.texture_image(img))
    tex.append(s.texture_checker(f(0.5, 2.0), tex[5], tex[4]))          # nested: image / marble under a checker
    mats = [s.material_lambertian(t) for t in tex]
    mats += [s.material_metal(f(0.3, 1), f(0.3, 1), f(0.3, 1), f(0, 1.2)) for _ in range(2)]
    mats += [s.material_dielectric(f(1.1, 2.0)), s.material_diffuse_light(tex[rs.randint(0, 3)]), s.material_isotropic(tex[1])]
    mat = lambda: mats[rs.randint(0, len(mats))]

    def prim():
        k = rs.randint(0, 7)
        c = (f(-4, 4), f(-1, 3), f(-4, 4))
        if k == 0:
            return s.sphere(c, f(0.2, 1.2), mat())
        if k == 1:
            c1 = (c[0] + f(0, 0.8), c[1] + f(0, 0.8), c[2] + f(0, 0.5))
            return s.moving_sphere(c, c1, 0.0, 1.0, f(0.2, 0.8), mat())
        if k == 2:
            return s.xy_rect(c[0], c[0] + f(0.3, 2), c[1], c[1] + f(0.3, 2), c[2], mat())
        if k == 3:
            return s.xz_rect(c[0], c[0] + f(0.3, 2), c[2], c[2] + f(0.3, 2), c[1], mat())
        if k == 4:
            return s.yz_rect(c[1], c[1] + f(0.3, 2), c[2], c[2] + f(0.3, 2), c[0], mat())
        if k == 5:
            p = np.array(c)
            return s.tri(p, p + rs.uniform(-1.5, 1.5, 3), p + rs.uniform(-1.5, 1.5, 3), mat())
        return s.cube(c, (c[0] + f(0.3, 1.5), c[1] + f(0.3, 1.5), c[2] + f(0.3, 1.5)), mat())

    def bvh(n):
        lst = s.list_create()
        items = [prim() for _ in range(n)]
        if n > 3:
            items.append(s.sphere((0.5, 0.5, 0.5), 0.4, mat()))
            items.append(s.sphere((0.5, 0.5, 0.5), 0.4, mat()))      # an exact duplicate: ties
        for it in items:
            s.list_add(lst, it)
        return s.bvh(lst, 0.0, 1.0, seed=int(rs.randint(0, 1 << 30)))

    world = []
    for _ in range(rs.randint(2, 5)):
        world.append(prim())
    world.append(bvh(rs.randint(1, 4)))
    world.append(bvh(rs.randint(8, 40)))
    world.append(s.translate(prim(), (f(-2, 2), f(-1, 1), f(-2, 2))))
    world.append(s.rotate_y(prim(), f(-60, 60)))
    world.append(s.translate(s.rotate_y(bvh(rs.randint(3, 12)), f(-40, 40)), (f(-2, 2), 0.0, f(-2, 2))))
    world.append(s.constant_medium(s.sphere((f(-2, 2), f(0, 2), f(-2, 2)), f(0.5, 1.5), mats[0]), f(0.2, 2.0), tex[rs.randint(0, 3)]))
    box = s.cube((0, 0, 0), (f(0.5, 1.5), f(0.5, 1.5), f(0.5, 1.5)), mats[1])
    world.append(s.constant_medium(s.translate(s.rotate_y(box, f(-30, 30)), (f(-2, 2), f(0, 1), f(-2, 2))), f(0.3, 3.0), tex[2]))
    nested = s.list_create()
    s.list_add(nested, prim())
    s.list_add(nested, prim())
    world.append(s.translate(nested, (f(-1, 1), 0.0, f(-1, 1))))       # Translate(List) is spliced
    rs.shuffle(world)
    for w in world:
        s.world_add(w)
    s.commit()
    return (f(0.2, 0.9), f(0.2, 0.9), f(0.2, 0.9))                     # background


def random_rays(seed, n):
    rs = np.random.RandomState(seed + 1000)
    o = rs.uniform(-6, 6, (n, 3))
    tgt = rs.uniform(-3, 3, (n, 3))
    d = (tgt - o) * rs.uniform(0.2, 3.0, (n, 1))                        # unnormalised directions, like the integrator's
    t = rs.uniform(0, 1, (n, 1))
    return np.concatenate([o, d, t], axis=1).astype(np.float32)
