"""Reference-shape topology tables (oracle; test infrastructure only).

Restates src/Grid/grid.py:196-229 of the reference (`reference_vertices`,
`reference_edges`, `reference_faces`) and `reference_face_edgenrs`
(src/Grid/grid.jl:261-275).  All indices here are 1-based like the reference.
"""

__all__ = ["REFSHAPES", "RefShape"]


class RefShape:
    def __init__(self, name, rdim, nvertices, edges, faces):
        self.name = name
        self.rdim = rdim
        self.nvertices = nvertices
        self.edges = tuple(edges)
        self.faces = tuple(faces)

    @property
    def face_edgenrs(self):
        """src/Grid/grid.jl:261-275"""
        out = []
        for face in self.faces:
            nrs = []
            for j in range(len(face)):
                v1, v2 = face[j], face[(j + 1) % len(face)]
                nr = None
                for k, e in enumerate(self.edges):
                    if e == (v1, v2) or e == (v2, v1):
                        nr = k + 1
                        break
                nrs.append(nr)
            out.append(tuple(nrs))
        return tuple(out)

    @property
    def facets(self):
        """Facets = codim-1 entities (src/Grid/grid.jl `reference_facets`)."""
        if self.rdim == 3:
            return self.faces
        if self.rdim == 2:
            return self.edges
        return tuple((v,) for v in range(1, self.nvertices + 1))

    def __repr__(self):
        return f"RefShape({self.name})"


REFSHAPES = {
    "line": RefShape("line", 1, 2, [(1, 2)], []),
    "triangle": RefShape("triangle", 2, 3, [(1, 2), (2, 3), (3, 1)], [(1, 2, 3)]),
    "quadrilateral": RefShape("quadrilateral", 2, 4, [(1, 2), (2, 3), (3, 4), (4, 1)], [(1, 2, 3, 4)]),
    "tetrahedron": RefShape(
        "tetrahedron", 3, 4,
        [(1, 2), (2, 3), (3, 1), (1, 4), (2, 4), (3, 4)],
        [(1, 3, 2), (1, 2, 4), (2, 3, 4), (1, 4, 3)],
    ),
    "hexahedron": RefShape(
        "hexahedron", 3, 8,
        [(1, 2), (2, 3), (3, 4), (4, 1), (5, 6), (6, 7), (7, 8), (8, 5), (1, 5), (2, 6), (3, 7), (4, 8)],
        [(1, 4, 3, 2), (1, 2, 6, 5), (2, 3, 7, 6), (3, 4, 8, 7), (1, 5, 8, 4), (5, 6, 7, 8)],
    ),
}
