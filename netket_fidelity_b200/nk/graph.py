"""Edge lists (NetKet `nk.graph.Chain`, `nk.graph.Grid`/`Hypercube`, `nk.graph.Graph`)."""


class Graph:
    def __init__(self, edges, n_nodes=None):
        self._edges = [(int(a), int(b)) for a, b in edges]
        self.n_nodes = n_nodes if n_nodes is not None else (max(max(e) for e in self._edges) + 1 if self._edges else 0)

    def edges(self):
        return list(self._edges)


class Chain(Graph):
    def __init__(self, length, pbc=True):
        e = [(i, i + 1) for i in range(length - 1)]
        if pbc and length > 2:
            e.append((length - 1, 0))
        super().__init__(e, length)


class Grid(Graph):
    def __init__(self, extent, pbc=True):
        Lx, Ly = extent
        e = set()
        for x in range(Lx):
            for y in range(Ly):
                i = x * Ly + y
                for dx, dy in ((1, 0), (0, 1)):
                    xx, yy = x + dx, y + dy
                    if pbc:
                        xx, yy = xx % Lx, yy % Ly
                    elif xx >= Lx or yy >= Ly:
                        continue
                    j = xx * Ly + yy
                    if i != j:
                        e.add((min(i, j), max(i, j)))
        super().__init__(sorted(e), Lx * Ly)


def Hypercube(length, n_dim=1, pbc=True):
    if n_dim == 1:
        return Chain(length, pbc)
    if n_dim == 2:
        return Grid([length, length], pbc)
    raise NotImplementedError
