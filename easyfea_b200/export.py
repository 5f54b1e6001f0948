"""Result hand-off of a sharded run without gathering it on one rank.

The reference writes, in MPI runs, one iteration pickle per rank holding the slices of the rank's own partition with the
nodes they cover (`_Simu.Save_Iter`, EasyFEA/Simulations/_simu.py:288-329, keys `__mpiLocalDofKeys` / `__mpiLocalNodes`), and
`Paraview.Save_simu` writes one `.vtu` piece per rank plus, on rank 0, the `.pvtu` descriptor and the `.pvd` timeline
(EasyFEA/Utilities/Paraview.py:26-145, 218-556; docs/howto/use_mpi.md "Export to ParaView").  Here the same wire formats are
produced from the device-resident fields of `easyfea_b200.staggered` / `transient` drivers: every rank writes the part it
OWNS (nodal fields over its owned nodes, cell fields over the elements of its own chunk), ghost cells / halo points of a
piece are flagged with `vtkGhostType`, and nothing travels between ranks.
"""
from __future__ import annotations

import base64
import os

import numpy as np

LOCAL_NODES_KEY = "__mpiLocalNodes"  # same names as the reference's pickles
LOCAL_DOFS_KEY = "__mpiLocalDofKeys"

# VTK cell type ids and the gmsh -> VTK node order (https://docs.vtk.org/en/latest/vtk_file_formats: linear cells share
# gmsh's order; the quadratic tetrahedron swaps its last two edge nodes, the 20/27-node hexahedra reorder edges and faces)
VTK_CELL = {"TRI3": 5, "TRI6": 22, "QUAD4": 9, "QUAD8": 23, "QUAD9": 28, "TETRA4": 10, "TETRA10": 24, "HEXA8": 12, "HEXA20": 25,
            "HEXA27": 29, "SEG2": 3, "SEG3": 21}
GMSH_TO_VTK = {"TETRA10": [0, 1, 2, 3, 4, 5, 6, 7, 9, 8],
               "HEXA20": [0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 13, 9, 16, 18, 19, 17, 10, 12, 14, 15],
               "HEXA27": [0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 13, 9, 16, 18, 19, 17, 10, 12, 14, 15, 22, 23, 21, 24, 20, 25, 26]}


def _host(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a)


# ---------------------------------------------------------------------------------------------------------
# iteration files: one per rank, slices of the owned dofs + the nodes they cover
# ---------------------------------------------------------------------------------------------------------
def save_iter(folder: str, Niter: int, fields: dict, part=None, scalars: dict = None) -> str:
    """Write iteration `Niter` of this rank: `fields` = {name: nodal array over `[owned | halo]` or over the owned dofs, any
    dof_n}; only the owned part is stored, with the global ids of the owned nodes.  `part`: `easyfea_b200.dist.Partition`
    (None: single GPU, every node owned).  `scalars`: convergence information etc., stored as they are.  Returns the path."""
    rank, world = (0, 1) if part is None else (part.rank, part.world)
    suffix = f"_rank{rank}" if world > 1 else ""
    path = os.path.join(folder, "Results", f"results{int(Niter)}{suffix}.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    out, keys = {}, []
    n_owned = None if part is None else part.n_owned
    for name, arr in fields.items():
        a = _host(arr)
        if part is not None:
            n_loc = part.n_local
            if a.shape[0] % n_loc == 0 and a.shape[0] != (a.shape[0] // n_loc) * n_owned:
                d = a.shape[0] // n_loc
                a = a.reshape(n_loc, d, *a.shape[1:])[:n_owned].reshape(n_owned * d, *a.shape[1:])
            elif a.shape[0] % max(n_owned, 1) != 0:
                raise ValueError(f"field {name}: length {a.shape[0]} is neither over the local nor over the owned nodes")
        out[name] = a
        keys.append(name)
    out[LOCAL_DOFS_KEY] = np.array(keys)
    out[LOCAL_NODES_KEY] = np.empty(0, dtype=np.int64) if part is None else np.asarray(part.nodes[: part.n_owned], dtype=np.int64)
    for k, v in (scalars or {}).items():
        out[f"scalar:{k}"] = np.asarray(v)
    np.savez(path, **out)
    return path


def load_iter(folder: str, Niter: int, Nn: int, world: int = 1) -> dict:
    """Read iteration `Niter` back as GLOBAL nodal arrays (what the reference rebuilds from the per-rank pickles): every rank's
    slices are placed at the nodes stored with them."""
    res, scal = {}, {}
    for r in range(world):
        suffix = f"_rank{r}" if world > 1 else ""
        d = np.load(os.path.join(folder, "Results", f"results{int(Niter)}{suffix}.npz"), allow_pickle=False)
        nodes = d[LOCAL_NODES_KEY]
        for name in d[LOCAL_DOFS_KEY]:
            a = d[str(name)]
            nodes_r = np.arange(Nn) if world == 1 else nodes  # a single rank stores whole fields
            dofn = a.shape[0] // max(nodes_r.size, 1) if nodes_r.size else 1
            full = res.setdefault(str(name), np.zeros((Nn * dofn,) + a.shape[1:], dtype=a.dtype))
            full.reshape(Nn, dofn, *a.shape[1:])[nodes_r] = a.reshape(nodes_r.size, dofn, *a.shape[1:])
        for k in d.files:
            if k.startswith("scalar:"):
                scal[k[7:]] = d[k]
    res["scalars"] = scal
    return res


# ---------------------------------------------------------------------------------------------------------
# ParaView: one .vtu piece per rank, .pvtu + .pvd on rank 0
# ---------------------------------------------------------------------------------------------------------
def _data_array(name, a, ncomp=None) -> str:
    a = np.ascontiguousarray(a)
    t = {"float64": "Float64", "float32": "Float32", "int64": "Int64", "int32": "Int32", "uint8": "UInt8"}[a.dtype.name]
    raw = a.tobytes()
    payload = base64.b64encode(np.uint64(len(raw)).tobytes() + raw).decode()
    nc = f' NumberOfComponents="{ncomp}"' if ncomp else ""
    return f'<DataArray type="{t}" Name="{name}"{nc} format="binary">{payload}</DataArray>\n'


def _vec3(a, n):
    """nodal field of dof_n components -> (n, 3) (ParaView wants 3-vectors) or (n,) for scalars"""
    a = _host(a).reshape(n, -1)
    if a.shape[1] == 1:
        return a[:, 0], None
    out = np.zeros((n, 3), dtype=a.dtype)
    out[:, : a.shape[1]] = a[:, :3]
    return out, 3


def save_vtu(folder: str, name: str, elemType: str, coords_local, connect_local, point_fields: dict = None, cell_fields: dict = None,
             part=None, time_index: int = 0) -> str:
    """Write this rank's piece `<name>_<time_index>[_rank<r>].vtu`: the local nodes `[owned | halo]` and ALL local elements (own
    chunk + ghosts), with `vtkGhostType` = 1 on halo points and ghost cells so that ParaView shows every entity once.
    `point_fields`: {name: array over the local nodes, dof_n components}; `cell_fields`: {name: array over the local elements}.
    Rank 0 also writes the `.pvtu` descriptor.  Returns the file a `.pvd` entry should reference."""
    rank, world = (0, 1) if part is None else (part.rank, part.world)
    os.makedirs(folder, exist_ok=True)
    X = _host(coords_local)
    conn = _host(connect_local).astype(np.int64)
    n_pts, (ne, nPe) = X.shape[0], conn.shape
    conn = conn[:, GMSH_TO_VTK.get(elemType, list(range(nPe)))]
    pts = np.zeros((n_pts, 3))
    pts[:, : X.shape[1]] = X[:, :3]
    base = f"{name}_{int(time_index)}"
    piece = f"{base}_rank{rank}.vtu" if world > 1 else f"{base}.vtu"
    pghost = np.zeros(n_pts, dtype=np.uint8)
    cghost = np.zeros(ne, dtype=np.uint8)
    if part is not None:
        pghost[part.n_owned:] = 1
        # a cell is shown by the rank that owns its lowest-owner node (that rank holds it for sure: it touches an owned node);
        # every other copy of it is a ghost cell.  Owners of the local nodes: this rank, then the halo segments.
        owner = np.full(n_pts, rank, dtype=np.int64)
        for i, q in enumerate(part.halo_ranks):
            owner[part.n_owned + int(part.halo_ptr[i]): part.n_owned + int(part.halo_ptr[i + 1])] = int(q)
        cghost[owner[_host(connect_local).astype(np.int64)].min(axis=1) != rank] = 1
    pdesc, cdesc = [], []
    with open(os.path.join(folder, piece), "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n')
        f.write(f'<UnstructuredGrid>\n<Piece NumberOfPoints="{n_pts}" NumberOfCells="{ne}">\n')
        f.write("<PointData>\n")
        for k, a in (point_fields or {}).items():
            v, nc = _vec3(a, n_pts)
            f.write(_data_array(k, v, nc))
            pdesc.append((k, v.dtype.name, nc))
        if world > 1:
            f.write(_data_array("vtkGhostType", pghost))
        f.write("</PointData>\n<CellData>\n")
        for k, a in (cell_fields or {}).items():
            v = _host(a).reshape(ne, -1)
            nc = v.shape[1] if v.shape[1] > 1 else None
            f.write(_data_array(k, v if nc else v[:, 0], nc))
            cdesc.append((k, v.dtype.name, nc))
        if world > 1:
            f.write(_data_array("vtkGhostType", cghost))
        f.write("</CellData>\n<Points>\n" + _data_array("Points", pts, 3) + "</Points>\n<Cells>\n")
        f.write(_data_array("connectivity", conn.ravel()))
        f.write(_data_array("offsets", (np.arange(ne, dtype=np.int64) + 1) * nPe))
        f.write(_data_array("types", np.full(ne, VTK_CELL[elemType], dtype=np.uint8)))
        f.write("</Cells>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    if world == 1:
        return piece
    pvtu = f"{base}.pvtu"
    if rank == 0:
        tname = {"float64": "Float64", "float32": "Float32", "int64": "Int64", "int32": "Int32", "uint8": "UInt8"}
        with open(os.path.join(folder, pvtu), "w") as f:
            f.write('<?xml version="1.0"?>\n<VTKFile type="PUnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n')
            f.write('<PUnstructuredGrid GhostLevel="1">\n<PPointData>\n')
            for k, dt, nc in pdesc:
                f.write(f'<PDataArray type="{tname[dt]}" Name="{k}"' + (f' NumberOfComponents="{nc}"' if nc else "") + "/>\n")
            f.write('<PDataArray type="UInt8" Name="vtkGhostType"/>\n</PPointData>\n<PCellData>\n')
            for k, dt, nc in cdesc:
                f.write(f'<PDataArray type="{tname[dt]}" Name="{k}"' + (f' NumberOfComponents="{nc}"' if nc else "") + "/>\n")
            f.write('<PDataArray type="UInt8" Name="vtkGhostType"/>\n</PCellData>\n')
            f.write('<PPoints>\n<PDataArray type="Float64" Name="Points" NumberOfComponents="3"/>\n</PPoints>\n')
            for r in range(world):
                f.write(f'<Piece Source="{base}_rank{r}.vtu"/>\n')
            f.write("</PUnstructuredGrid>\n</VTKFile>\n")
    return pvtu


def save_pvd(folder: str, name: str, files, times=None) -> str:
    """the `.pvd` timeline (rank 0): `files[i]` = what `save_vtu` returned for time step i"""
    path = os.path.join(folder, f"{name}.pvd")
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1" byte_order="LittleEndian">\n<Collection>\n')
        for i, fn in enumerate(files):
            t = i if times is None else times[i]
            f.write(f'<DataSet timestep="{t}" group="" part="1" file="{fn}"/>\n')
        f.write("</Collection>\n</VTKFile>\n")
    return path
