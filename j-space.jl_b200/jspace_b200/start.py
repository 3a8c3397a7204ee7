"""Start_J_Space(parameters, config) for the dynamics stage: reads the reference's
Parameters.toml / Config.toml (src/Start.jl:4-100), runs the GPU dynamics and writes the four
files of src/Start.jl:101-116,149-161 — Dinamica.csv, n_cell_alive.csv, DriverList.csv,
CA_subpop.csv — in the reference's CSV.jl formats.  The later stages of Start_J_Space (sampling,
trees, molecular evolution, bulk, ART: src/Start.jl:163-403) are out of scope.
"""
from __future__ import annotations

import os
import tomllib
import uuid

import numpy as np

from . import _lib as L
from .api import PhiloxSeed, simulate_evolution, spatial_graph


def julia_float(x: float) -> str:
    """Julia's shortest round-trip printing of a Float64 (what CSV.jl writes)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    r = repr(x)
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        if "." not in mant:
            mant += ".0"
        return f"{mant}e{int(exp)}"
    if x != 0 and (abs(x) >= 1e16 or abs(x) < 1e-5):
        # Julia switches to exponent form outside [1e-5, 1e16)
        digits = r.replace("-", "").replace(".", "").lstrip("0")
        exp = int(np.floor(np.log10(abs(x))))
        mant = digits[0] + "." + (digits[1:] or "0")
        return ("-" if x < 0 else "") + f"{mant}e{exp}"
    return r


def julia_repr(v) -> str:
    """repr of a genotype as CSV.jl prints the `Driver` column: 1, [1, 143], driver_1,
    Any["driver_1", "driver_2"]."""
    if isinstance(v, list):
        if v and isinstance(v[0], str):
            return "Any[" + ", ".join('"' + x + '"' for x in v) + "]"
        return "[" + ", ".join(str(x) for x in v) + "]"
    return str(v)


def csv_field(s: str) -> str:
    """CSV.jl quoting: fields containing the delimiter, a quote or a newline are quoted, quotes doubled."""
    if any(c in s for c in ',"\n\r'):
        return '"' + s.replace('"', '""') + '"'
    return s


def cell_uuid(ident: int) -> str:
    """Deterministic UUID text for a sequential cell id (the reference's ids are uuid1 values)."""
    return str(uuid.UUID(int=int(ident)))


def write_dinamica(path, df):
    with open(path, "w") as f:
        f.write("Event,Time,Cell,Notes\n")
        ev, tm, cell = df.Event, df.Time, df.Cell
        for i in range(len(df)):
            notes = df.notes(i)
            if notes is None:
                txt = "UndefInitializer()"
            elif len(notes) == 1:
                txt = f'UUID[UUID("{cell_uuid(notes[0])}")]'
            else:
                second = julia_repr(notes[1]) if not isinstance(notes[1], str) else '"' + notes[1] + '"'
                txt = f'Any[UUID("{cell_uuid(notes[0])}"), {second}]'
            f.write(f"{ev[i]},{julia_float(tm[i])},{cell_uuid(cell[i])},{csv_field(txt)}\n")


def write_vector_table(path, values, formatter=str):
    """CSV.write(path, Tables.table(v), header=false): one value per line."""
    with open(path, "w") as f:
        for v in values:
            f.write(csv_field(formatter(v)) + "\n")


def write_driver_list(path, genotypes, alpha, header: bool):
    with open(path, "w") as f:
        if header:
            f.write("Driver,Fitness\n")
        for g, a in zip(genotypes, alpha):
            f.write(f"{csv_field(julia_repr(g))},{julia_float(a)}\n")


def Start_J_Space(parameters: str, config: str, *, device: int = 0):
    with open(parameters, "rb") as f:
        par = tomllib.load(f)
    with open(config, "rb") as f:
        conf = tomllib.load(f)
    cfg = conf["Config"][0]
    seed = PhiloxSeed(int(cfg["seed"]))                                   # src/Start.jl:12-13
    save = cfg["path_to_save_files"]
    os.makedirs(save, exist_ok=True)
    start_cell = par["Graph"][0]["N_starting_cells"]                      # :31
    if cfg["Generate_graph"] == 0:                                        # :32-34
        g_meta = spatial_graph(cfg["Path_to_Graph"], seed, n_cell=start_cell)
    else:                                                                 # :36-44
        gp = par["Graph"][0]
        g_meta = spatial_graph(gp["row"], gp["col"], seed, dim=gp["dim"], n_cell=start_cell)
    plot = conf["FileOutputPlot"][0]
    tos = plot["Time_of_sampling"] if plot["Graph_configuration"] == 1 else []   # :50-55
    dyn = par["Dynamic"][0]
    kw = dict(Time_of_sampling=tos, t_bottleneck=dyn["t_bottleneck"], ratio_bottleneck=dyn["ratio_bottleneck"],
              device=device)
    if cfg["Tree_Driver_Configure"] != 1:                                 # :66-83
        out = simulate_evolution(g_meta, dyn["Max_time"], dyn["rate_birth"], dyn["rate_death"], dyn["rate_migration"],
                                 dyn["drive_mut_rate"], dyn["average_driver_mut_rate"], dyn["std_driver_mut_rate"],
                                 dyn["Model"], seed, **kw)
    else:                                                                 # :84-99
        out = simulate_evolution(g_meta, dyn["Max_time"], dyn["rate_death"], dyn["rate_migration"],
                                 dyn["drive_mut_rate"], dyn["Model"], cfg["edgelist_treedriver"],
                                 cfg["driver_birth_rates"], seed, **kw)
    df, G, n_cell_alive, set_mut, gs_conf, ca_subpop, alpha = out
    write_dinamica(os.path.join(save, "Dinamica.csv"), df)                                  # :101-103
    write_vector_table(os.path.join(save, "n_cell_alive.csv"), n_cell_alive)                # :105-107
    write_driver_list(os.path.join(save, "DriverList.csv"), set_mut, alpha, header=True)    # :109-112
    write_vector_table(os.path.join(save, "CA_subpop.csv"), ca_subpop, julia_repr)          # :114-116
    if conf["OutputGT"][0]["Driver_list"] == 1:                                             # :149-161
        write_driver_list(os.path.join(save, "DriverList.csv"), set_mut, alpha, header=False)
    return out
