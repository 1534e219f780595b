"""The reference's own scenarios (fiksi/src/tests/*.rs, the lib.rs doc example, the
fiksi_svg_tests example and the criterion bench shape), written against the minimal System
interface shared by ``oracle.System`` and ``fiksi_b200.System``.  Each builder returns a dict of
handles; citations give the reference lines the construction order is copied from (the order
matters: it decides variable indices, component order and the perturbation stream)."""
import math

RAD = math.pi / 180.0  # f64::to_radians multiplies by PI / 180.0


def deg(x):
    return x * RAD


def coincident_points(S):  # tests/basic.rs:9-34
    s = S()
    p0, p1 = s.add_point(0., 0.), s.add_point(1., 0.5)
    c = s.point_point_coincidence(p0, p1)
    return dict(s=s, points=[p0, p1], constraints=[c])


def underconstrained_triangle(S):  # tests/basic.rs:36-52
    s = S()
    p = [s.add_point(0., 0.), s.add_point(1., 0.5), s.add_point(2., 1.)]
    c = [s.point_point_point_angle(p[0], p[1], p[2], deg(40.)),
         s.point_point_point_angle(p[1], p[2], p[0], deg(80.))]
    return dict(s=s, points=p, constraints=c)


def overconstrained_triangle_line_incidence(S):  # tests/basic.rs:54-87
    s = S()
    p = [s.add_point(0., 0.), s.add_point(1., 0.5), s.add_point(2., 1.), s.add_point(3., 1.5)]
    line0 = s.add_line(p[2], p[3])
    c = [s.point_point_point_angle(p[0], p[1], p[2], deg(40.)),
         s.point_point_point_angle(p[1], p[2], p[0], deg(80.)),
         s.point_point_point_angle(p[2], p[0], p[1], deg(100.)),
         s.point_line_incidence(p[1], line0)]
    return dict(s=s, points=p, constraints=c)


def overconstrained(S):  # tests/basic.rs:89-112 (solved instead of analysed)
    s = S()
    p = [s.add_point(0.123, 0.1), s.add_point(1.2, 0.), s.add_point(-0.5, 1.1), s.add_point(1.599, 1.2)]
    c = [s.point_point_distance(p[0], p[1], 1.), s.point_point_distance(p[0], p[2], 1.5),
         s.point_point_distance(p[1], p[3], 1.7), s.point_point_distance(p[2], p[3], 1.2),
         s.point_point_distance(p[1], p[2], 2.), s.point_point_distance(p[0], p[3], 5.)]
    return dict(s=s, points=p, constraints=c)


def triangle_inscribed_circle(S):  # tests/basic.rs:114-149
    s = S()
    p = [s.add_point(0., 0.), s.add_point(1., 0.5), s.add_point(1.5, 1.), s.add_point(2.8, 1.5)]
    c = [s.point_point_distance(p[0], p[1], 1.), s.point_point_distance(p[0], p[2], 1.),
         s.point_point_distance(p[1], p[2], 1.)]
    l0, l1, l2 = s.add_line(p[0], p[1]), s.add_line(p[0], p[2]), s.add_line(p[1], p[2])
    radius = s.add_length(1.)
    circle = s.add_circle(p[3], radius)
    c += [s.line_circle_tangency(l0, circle), s.line_circle_tangency(l1, circle),
          s.line_circle_tangency(l2, circle)]
    return dict(s=s, points=p, constraints=c, radius=radius)


def two_connected_components(S):  # tests/basic.rs:151-170
    s = S()
    p = [s.add_point(0.123, 0.1), s.add_point(1.2, 0.), s.add_point(-0.5, 1.1), s.add_point(1.599, 1.2)]
    c = [s.point_point_distance(p[0], p[1], 1.), s.point_point_distance(p[2], p[3], 1.2)]
    return dict(s=s, points=p, constraints=c)


def single_triangle(S, fixed=None):  # tests/triangles.rs:8-37, tests/fixed.rs:9-43
    s = S()
    p = [s.add_point(0., 0.), s.add_point(1., 0.5), s.add_point(2., 1.)]
    if fixed is not None:
        s.fix(p[fixed])
    c = [s.point_point_distance(p[0], p[1], 1.), s.point_point_distance(p[0], p[2], 1.),
         s.point_point_distance(p[1], p[2], 1.)]
    return dict(s=s, points=p, constraints=c)


def connected_triangles(S):  # tests/triangles.rs:39-68
    s = S()
    p = [s.add_point(float(i), 0.5 * i) for i in range(6)]
    c = [s.point_point_point_angle(p[5], p[0], p[1], deg(-135.)),
         s.point_point_point_angle(p[1], p[2], p[3], deg(-120.)),
         s.point_point_point_angle(p[3], p[4], p[5], deg(-115.)),
         s.point_point_distance(p[0], p[1], 7.), s.point_point_distance(p[1], p[2], 5.),
         s.point_point_distance(p[2], p[3], 9.), s.point_point_distance(p[3], p[4], 8.),
         s.point_point_distance(p[4], p[5], 6.), s.point_point_distance(p[5], p[0], 7.)]
    return dict(s=s, points=p, constraints=c)


def hinged_triangles_test(S):  # tests/triangles.rs:70-104
    s = S()
    p = [s.add_point(0.5, 0.), s.add_point(1.1, 0.5), s.add_point(2.1, 1.), s.add_point(3.1, 1.5),
         s.add_point(4.1, 2.), s.add_point(5.1, 2.5), s.add_point(6.1, 3.)]
    c = []
    for a, b in ((1, 2), (3, 4), (5, 6)):
        c += [s.point_point_distance(p[0], p[a], 1.), s.point_point_distance(p[0], p[b], 1.),
              s.point_point_distance(p[a], p[b], 1.)]
    return dict(s=s, points=p, constraints=c)


def collinear_points(S):  # tests/singular.rs:19-40
    s = S()
    p = [s.add_point(0., 0.), s.add_point(3., 0.), s.add_point(6., 0.)]
    c = [s.point_point_distance(p[0], p[1], 1.), s.point_point_distance(p[0], p[2], 1.),
         s.point_point_distance(p[1], p[2], 1.)]
    return dict(s=s, points=p, constraints=c)


def fixed_point_and_circle_center_incidence(S):  # tests/fixed.rs:45-80
    s = S()
    p0, center = s.add_point(0., 0.), s.add_point(4., 3.)
    radius = s.add_length(1.)
    circle = s.add_circle(center, radius)
    s.fix(p0)
    s.fix(center)
    c = [s.point_circle_incidence(p0, circle)]
    return dict(s=s, points=[p0, center], constraints=c, radius=radius)


def fixed_with_coincidence(S):  # tests/fixed.rs:82-127
    s = S()
    p = [s.add_point(0., 0.), s.add_point(1., 0.5), s.add_point(2., 1.), s.add_point(5., 5.)]
    s.fix(p[3])
    c = [s.point_point_distance(p[0], p[1], 1.), s.point_point_distance(p[1], p[2], 1.),
         s.point_point_coincidence(p[2], p[3])]
    return dict(s=s, points=p, constraints=c)


def large_order_of_magnitude(S):  # tests/magnitude.rs:8-36
    F = 1e20
    s = S()
    p = [s.add_point(1.5 * F, 6.5 * F), s.add_point(3.2 * F, 0.8 * F), s.add_point(2.2 * F, -1.5 * F)]
    c = [s.point_point_distance(p[0], p[1], 5. * F), s.point_point_distance(p[0], p[2], 3. * F),
         s.point_point_distance(p[1], p[2], 4. * F)]
    return dict(s=s, points=p, constraints=c, factor=F)


def _four_point_frame(S, F):
    s = S()
    p = [s.add_point(1.5 * F, 6.5 * F), s.add_point(3.2 * F, 0.8 * F), s.add_point(2.2 * F, -1.5 * F),
         s.add_point(1.2 * F, 0.5 * F)]
    c = [s.point_point_distance(p[0], p[1], 5. * F), s.point_point_distance(p[1], p[2], 4. * F),
         s.point_point_distance(p[2], p[3], 3. * F), s.point_point_distance(p[3], p[1], 1. * F)]
    l0, l1 = s.add_line(p[0], p[1]), s.add_line(p[2], p[3])
    return s, p, c, l0, l1


def distance_and_angle(S):  # tests/magnitude.rs:38-86
    F = 1e10
    s, p, c, l0, l1 = _four_point_frame(S, F)
    c.append(s.line_line_angle(l0, l1, deg(30.)))
    return dict(s=s, points=p, constraints=c, factor=F)


def metric_and_singular(S):  # tests/magnitude.rs:88-137
    F = 1e7
    s, p, c, l0, l1 = _four_point_frame(S, F)
    c.append(s.line_line_parallelism(l0, l1))
    return dict(s=s, points=p, constraints=c, factor=F)


def near_degenerate_isosceles_triangle(S):  # tests/magnitude.rs:139-166
    F = 1e13
    s = S()
    p = [s.add_point(1.5 * F, 6.5 * F), s.add_point(3.2 * F, 0.8 * F), s.add_point(2.2, -1.5)]
    c = [s.point_point_distance(p[0], p[1], 4. * F + 1.), s.point_point_distance(p[1], p[2], 4. * F + 1.),
         s.point_point_distance(p[0], p[2], 1.)]
    return dict(s=s, points=p, constraints=c, factor=F)


def lib_doc_example(S):  # fiksi/src/lib.rs:19-33
    s = S()
    p = [s.add_point(1., 0.), s.add_point(0.8, 1.), s.add_point(1.1, 2.)]
    c = [s.point_point_distance(p[1], p[2], 5.), s.point_point_point_angle(p[0], p[1], p[2], deg(10.)),
         s.point_point_point_angle(p[1], p[2], p[0], deg(60.))]
    return dict(s=s, points=p, constraints=c)


def circle_triangle_line(S, scale=1.0, noise=None):
    """examples/fiksi_svg_tests/src/main.rs:9-45 — BASELINE config 1 (and the topology of config 4).
    `scale`/`noise` implement SURVEY App. D's C4 perturbation: coords s*(c + 0.05*|c|*U)."""
    base = [(10., 0.), (20., 10.), (30., -10.), (-40., -50.), (40., -50.)]
    if noise is None:
        noise = [0.0] * 10
    s = S()
    p = []
    for i, (x, y) in enumerate(base):
        p.append(s.add_point(scale * (x + 0.05 * abs(x) * noise[2 * i]), scale * (y + 0.05 * abs(y) * noise[2 * i + 1])))
    c = [s.point_point_point_angle(p[0], p[1], p[2], deg(40.)),
         s.point_point_point_angle(p[1], p[2], p[0], deg(70.)),
         s.point_point_distance(p[0], p[1], 70. * scale)]
    side1, side2, side3 = s.add_line(p[0], p[1]), s.add_line(p[1], p[2]), s.add_line(p[0], p[2])
    radius = s.add_length(5. * scale)
    circle = s.add_circle(p[2], radius)
    c.append(s.line_circle_tangency(side1, circle))
    line = s.add_line(p[3], p[4])
    c.append(s.line_line_angle(side3, line, deg(-90.)))
    c.append(s.point_line_incidence(p[2], line))
    c.append(s.point_point_distance(p[2], p[3], 40. * scale))
    c.append(s.point_point_distance(p[3], p[4], 80. * scale))
    return dict(s=s, points=p, constraints=c, radius=radius)


def hinged_triangles_bench(S, n):  # fiksi/benches/fiksi_bench.rs:15-40
    s = S()
    hinge = s.add_point(0., 0.)
    p, c = [hinge], []
    for k in range(n):
        p1, p2 = s.add_point(-1., float(k)), s.add_point(1., float(k))
        p += [p1, p2]
        c += [s.point_point_distance(hinge, p1, 2.), s.point_point_distance(hinge, p2, 2.),
              s.point_point_distance(p1, p2, 3.)]
    return dict(s=s, points=p, constraints=c)


def stale_component_quirk(S):
    """SURVEY F7 (fiksi/src/graph.rs:178-225): two triangles built separately, bridged, then one more
    distance inside the smaller triangle's stale points and one from a stale point to a fresh one."""
    s = S()
    a = [s.add_point(0., 0.), s.add_point(1., 0.1), s.add_point(0.4, 0.9), s.add_point(1.5, 1.2)]
    b = [s.add_point(5., 0.), s.add_point(6., 0.2), s.add_point(5.5, 1.1)]
    fresh = s.add_point(8., 8.)
    c = []
    # bigger component A: 4 points
    c += [s.point_point_distance(a[0], a[1], 1.), s.point_point_distance(a[1], a[2], 1.),
          s.point_point_distance(a[0], a[2], 1.), s.point_point_distance(a[2], a[3], 1.)]
    # smaller component B: 3 points
    c += [s.point_point_distance(b[0], b[1], 1.), s.point_point_distance(b[1], b[2], 1.)]
    # bridge A-B through a[3], b[0]: b[1], b[2] keep a stale component index
    c.append(s.point_point_distance(a[3], b[0], 3.))
    # both stale: opens a component without elements -> never enforced
    c.append(s.point_point_distance(b[1], b[2], 2.5))
    # stale + fresh: component {fresh} only, b[2] held fixed at its pre-solve value
    c.append(s.point_point_distance(b[2], fresh, 2.))
    return dict(s=s, points=a + b + [fresh], constraints=c)


REFERENCE_SOLVED = {  # scenario -> reference assertion kind
    "coincident_points": coincident_points,
    "underconstrained_triangle": underconstrained_triangle,
    "triangle_inscribed_circle": triangle_inscribed_circle,
    "two_connected_components": two_connected_components,
    "single_triangle": single_triangle,
    "connected_triangles": connected_triangles,
    "hinged_triangles": hinged_triangles_test,
    "collinear_points": collinear_points,
    "fixed_with_coincidence": fixed_with_coincidence,
}

ALL = dict(REFERENCE_SOLVED)
ALL.update({
    # examples without a reference assertion
    "lib_doc_example": lib_doc_example,
    "circle_triangle_line": circle_triangle_line,
    "overconstrained_triangle_line_incidence": overconstrained_triangle_line_incidence,
    "overconstrained": overconstrained,
    "single_triangle_fixed_p1": lambda S: single_triangle(S, fixed=1),
    "fixed_point_and_circle_center_incidence": fixed_point_and_circle_center_incidence,
    "large_order_of_magnitude": large_order_of_magnitude,
    "distance_and_angle": distance_and_angle,
    "metric_and_singular": metric_and_singular,
    "near_degenerate_isosceles_triangle": near_degenerate_isosceles_triangle,
    "hinged_triangles_bench_4": lambda S: hinged_triangles_bench(S, 4),
    "hinged_triangles_bench_16": lambda S: hinged_triangles_bench(S, 16),
    "stale_component_quirk": stale_component_quirk,
})


def rms(vals):
    vals = list(vals)
    return math.sqrt(sum(v * v for v in vals) / len(vals))
