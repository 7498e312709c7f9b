"""Write a .prm file for cracks_b200_run from the parameter dictionary a golden fixture carries (the keys are
the reference's own, cracks.cc:1290-1480), so the command-line tests run the same input the reference ran."""

SECTIONS = {
    "Global parameters": ["Global pre-refinement steps", "Local pre-refinement steps", "Adaptive refinement cycles",
                          "Max No of timesteps", "Timestep size", "Timestep size to switch to", "Switch timestep after steps",
                          "outer solver", "test case", "ref strategy", "value phase field for refinement"],
    "Problem dependent parameters": ["K reg", "Eps reg", "Gamma penalization", "Pressure", "Fracture toughness G_c",
                                     "Poisson ratio nu", "E modulus", "Lame mu", "Lame lambda"],
    "Solver parameters": ["Use Direct Inner Solver", "Newton lower bound", "Newton maximum steps", "Upper Newton rho",
                          "Line search maximum steps", "Line search damping", "Decompose stress in rhs",
                          "Decompose stress in matrix"],
}


def write_prm(path, prm, dim, output_dir, exact=None, **overrides):
    """overrides: keyword = key with blanks replaced by underscores; `exact`: dict with the keys as they are
    (for the ones that contain a hyphen)"""
    p = dict(prm)
    p.update({k.replace("_", " "): v for k, v in overrides.items()})
    p.update(exact or {})
    unknown = [k for k in p if not any(k in keys for keys in SECTIONS.values()) and k != "Output filename"]
    assert not unknown, unknown
    lines = []
    for sec, keys in SECTIONS.items():
        lines.append("subsection " + sec)
        if sec == "Global parameters":
            lines += ["  set Dimension = %d" % dim, "  set Output directory = %s" % output_dir]
        lines += ["  set %s = %s" % (k, p[k]) for k in keys if k in p]
        lines.append("end")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def read_statistics(path):
    return [l.split() for l in open(path) if not l.startswith("#")]
