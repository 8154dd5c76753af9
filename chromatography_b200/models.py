"""Host-side mirror of the reference's model structs for the hot path.

These carry parameters only; all arithmetic runs in libchrom_cuda.so.  Names, argument order,
validation and error texts follow chrom-rs:
  TemporalInjection   src/models/injection.rs:92-178, 349-399
  LangmuirSingle      src/models/langmuir_single.rs:154-198, 235-276
  SpeciesParams       src/models/langmuir_multi.rs:136-204
  LangmuirMulti       src/models/langmuir_multi.rs:287-316, 342-410, 477-495
Derived fields (dz, fe, ue, stationary_fraction) are plain attributes because the reference
deserialises them verbatim from model files (src/config/model.rs:111-148): `fe: 1.5` in YAML is
not the same double as (1-0.4)/0.4.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from decimal import Decimal

from . import _ffi


def rust_f64(x: float) -> str:
    """`f64::to_string()` / `{}`: shortest round-trip digits, never exponent notation, no '.0'."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    s = format(Decimal(repr(x)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s


@dataclass(frozen=True)
class TemporalInjection:
    """Inlet profile C(z=0, t).  `Custom` closures are not representable across the C ABI."""
    kind: int = _ffi.INJ_NONE
    p0: float = 0.0
    p1: float = 0.0
    p2: float = 0.0

    @staticmethod
    def none() -> "TemporalInjection":
        return TemporalInjection(_ffi.INJ_NONE)

    @staticmethod
    def dirac(time: float, amount: float) -> "TemporalInjection":
        return TemporalInjection(_ffi.INJ_DIRAC, float(time), float(amount))

    @staticmethod
    def gaussian(center: float, width: float, peak_concentration: float) -> "TemporalInjection":
        return TemporalInjection(_ffi.INJ_GAUSSIAN, float(center), float(width), float(peak_concentration))

    @staticmethod
    def rectangle(start: float, end: float, concentration: float) -> "TemporalInjection":
        if not end > start:
            raise ValueError("Rectangle end must be > start")
        return TemporalInjection(_ffi.INJ_RECTANGLE, float(start), float(end), float(concentration))

    def to_map(self) -> dict:
        """The serde snapshot shape (injection.rs:54-72)."""
        if self.kind == _ffi.INJ_DIRAC:
            return {"type": "Dirac", "time": self.p0, "amount": self.p1}
        if self.kind == _ffi.INJ_GAUSSIAN:
            return {"type": "Gaussian", "center": self.p0, "width": self.p1, "peak_concentration": self.p2}
        if self.kind == _ffi.INJ_RECTANGLE:
            return {"type": "Rectangle", "start": self.p0, "end": self.p1, "concentration": self.p2}
        return {"type": "None"}

    @staticmethod
    def from_map(m: dict) -> "TemporalInjection":
        t = m.get("type", "None")
        if t == "Dirac":
            return TemporalInjection.dirac(m["time"], m["amount"])
        if t == "Gaussian":
            return TemporalInjection.gaussian(m["center"], m["width"], m["peak_concentration"])
        if t == "Rectangle":
            return TemporalInjection.rectangle(m["start"], m["end"], m["concentration"])
        if t == "None":
            return TemporalInjection.none()
        raise ValueError(f"unknown injection type '{t}' (Custom closures cannot cross the C ABI)")


@dataclass
class LangmuirSingle:
    lambda_: float
    langmuir_k: float
    port_number: float
    column_length: float
    n_points: int
    dz: float
    fe: float
    ue: float
    injection: TemporalInjection

    @staticmethod
    def new(lambda_, langmuir_k, port_number, porosity, velocity, column_length, spatial_points,
            injection) -> "LangmuirSingle":
        if not (porosity > 0.0 and porosity < 1.0):
            raise ValueError(f"Porosity must be in ]0,1[, got {rust_f64(porosity)}")
        if not column_length > 0.0:
            raise ValueError(f"Column length must be positive, got {rust_f64(column_length)}")
        if not spatial_points >= 2:
            raise ValueError(f"Need at least 2 spatial points, got {spatial_points}")
        fe = (1.0 - porosity) / porosity             # langmuir_single.rs:261
        ue = velocity / porosity                     # :262
        dz = column_length / float(spatial_points)   # :263
        return LangmuirSingle(float(lambda_), float(langmuir_k), float(port_number), float(column_length),
                              int(spatial_points), dz, fe, ue, injection)

    def points(self) -> int:
        return self.n_points

    def spatial_points(self) -> int:
        return self.n_points

    def set_injection(self, injection: TemporalInjection) -> None:
        self.injection = injection

    def name(self) -> str:
        return "Langmuir single specie with temporal injection"

    def setup_initial_state(self):
        import numpy as np
        return np.zeros(self.n_points)

    def to_map(self) -> dict:
        """Externally tagged serde shape (src/config/model.rs:111-148)."""
        return {"LangmuirSingle": {"lambda": self.lambda_, "langmuir_k": self.langmuir_k,
                                   "port_number": self.port_number, "column_length": self.column_length,
                                   "n_points": self.n_points, "dz": self.dz, "fe": self.fe, "ue": self.ue,
                                   "injection": self.injection.to_map()}}

    @staticmethod
    def from_map(m: dict) -> "LangmuirSingle":
        b = m["LangmuirSingle"] if "LangmuirSingle" in m else m
        return LangmuirSingle(b["lambda"], b["langmuir_k"], b["port_number"], b["column_length"], b["n_points"],
                              b["dz"], b["fe"], b["ue"], TemporalInjection.from_map(b.get("injection", {})))


@dataclass
class SpeciesParams:
    name: str
    lambda_: float
    langmuir_k: float
    port_number: int
    injection: TemporalInjection

    @staticmethod
    def new(name, lambda_, langmuir_k, port_number, injection) -> "SpeciesParams":
        return SpeciesParams(str(name), float(lambda_), float(langmuir_k), int(port_number), injection)

    def validate(self) -> None:
        if self.lambda_ < 0.0:
            raise ValueError(f"Species '{self.name}' : lambda must be >= 0, got {rust_f64(self.lambda_)}")
        if self.langmuir_k <= 0.0:
            raise ValueError(f"Species '{self.name}' : Langmuir K̃ must be > 0, got {rust_f64(self.langmuir_k)}")
        if self.port_number == 0:
            raise ValueError(f"Species '{self.name}' : Port number must be > 0 (competitiveness), "
                             f"got {self.port_number}")


@dataclass
class LangmuirMulti:
    species: list
    n_points: int
    porosity: float
    velocity: float
    column_length: float
    dz: float
    fe: float
    ue: float
    stationary_fraction: float
    _names: set = field(default_factory=set, repr=False)

    @staticmethod
    def new(species, n_points, porosity, velocity, column_length) -> "LangmuirMulti":
        species = list(species)
        if not species:
            raise ValueError("Langmuir multi species must have at least one species.")
        for sp in species:
            sp.validate()
        seen = set()
        for sp in species:
            if sp.name in seen:
                raise ValueError(f"Duplicate speciees name '{sp.name}' in intial species list")
            seen.add(sp.name)
        if n_points < 2:
            raise ValueError(f"number of points must have at least two points, got {n_points}.")
        if porosity <= 0.0 or porosity >= 1.0:
            raise ValueError(f"porosity must be in ]0, 1[, got {rust_f64(porosity)}.")
        if velocity <= 0.0:
            raise ValueError(f"velocity must be strictly positive, got {rust_f64(velocity)}.")
        if column_length <= 0.0:
            raise ValueError(f"column length must be strictly positive, got {rust_f64(column_length)}.")
        dz = column_length / float(n_points)   # langmuir_multi.rs:394
        fe = (1.0 - porosity) / porosity       # :395
        ue = velocity / porosity               # :396
        sf = 1.0 - porosity                    # :397
        return LangmuirMulti(species, int(n_points), float(porosity), float(velocity), float(column_length), dz, fe,
                             ue, sf, seen)

    def add_species(self, sp: SpeciesParams) -> None:
        sp.validate()
        if sp.name in {s.name for s in self.species}:
            raise ValueError(f"Species '{sp.name}' already exists in this model")
        self.species.append(sp)

    def n_species(self) -> int:
        return len(self.species)

    def points(self) -> int:
        return self.n_points

    def species_names(self) -> list:
        return [s.name for s in self.species]

    def name(self) -> str:
        return "Langmuir Multi-Species"

    def setup_initial_state(self):
        import numpy as np
        return np.zeros((self.n_points, self.n_species()))

    def to_map(self) -> dict:
        return {"LangmuirMulti": {
            "species": [{"name": s.name, "lambda": s.lambda_, "langmuir_k": s.langmuir_k,
                         "port_number": s.port_number, "injection": s.injection.to_map()} for s in self.species],
            "n_points": self.n_points, "porosity": self.porosity, "velocity": self.velocity,
            "column_length": self.column_length, "dz": self.dz, "fe": self.fe, "ue": self.ue,
            "stationary_fraction": self.stationary_fraction}}

    @staticmethod
    def from_map(m: dict) -> "LangmuirMulti":
        b = m["LangmuirMulti"] if "LangmuirMulti" in m else m
        sp = [SpeciesParams(s["name"], s["lambda"], s["langmuir_k"], int(s["port_number"]),
                            TemporalInjection.from_map(s.get("injection", {}))) for s in b["species"]]
        return LangmuirMulti(sp, b["n_points"], b["porosity"], b["velocity"], b["column_length"], b["dz"], b["fe"],
                             b["ue"], b["stationary_fraction"])
