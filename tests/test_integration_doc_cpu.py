"""INTEGRATION.md's C snippets are type-checked against include/beluga_b200.h: each ```c block is wrapped in a function that
declares the variables the prose takes for granted and compiled with gcc -fsyntax-only, so the document cannot drift away
from the ABI it describes."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROLOGUES = [
    # 2. The C ABI directly
    "int block0(const int8_t* cells, int width, int height, const double* mean_xytheta, const double* cov3x3,\n"
    "           const double* pose_cos_sin_x_y, const double* points_xy, uint64_t n_points) {\n",
    # 4. one process, several GPUs
    "int block1(const int8_t* cells, int width, int height, const double* mean_xytheta, const double* cov3x3, const double* pose,\n"
    "           const double* points_xy, uint64_t n_points, double* states, double* weights, uint64_t capacity) {\n"
    "  bb200_likelihood_field_param lfm = { 2.0, 100.0, 0.5, 0.5, 0.2, 0, 0 };\n"
    "  bb200_occupancy_grid grid = { cells, width, height, 0.05, {1, 0, 0, 0} };\n"
    "  bb200_update_result r;\n",
    # 4. one process per GPU
    "int block2(int world, int rank, int local_rank, const double* pose, const double* points_xy, uint64_t n_points) {\n"
    "  bb200_amcl_param p = { .min_particles = 4000000, .max_particles = 4000000 };\n"
    "  bb200_motion_param m = { BB200_MOTION_DIFFERENTIAL, 0.1, 0.05, 0.1, 0.05, 0.0, 0.01 };\n"
    "  bb200_amcl* a; bb200_update_result r;\n",
]


def test_c_snippets_compile_against_the_header(tmp_path):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```c\n(.*?)```", text, flags=re.S)
    assert len(blocks) == len(PROLOGUES), "INTEGRATION.md gained or lost a C snippet: give it a prologue here"
    src = "#include <stdio.h>\n#include <stdint.h>\n#include <beluga_b200.h>\n\n"
    for prologue, block in zip(PROLOGUES, blocks):
        body = "\n".join(line for line in block.splitlines() if not line.startswith("#include"))
        src += prologue + body + "\n  return 0;\n}\n\n"
    path = tmp_path / "integration_snippets.c"
    path.write_text(src)
    out = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-Wno-unused-but-set-variable",
                          "-Wno-missing-field-initializers", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(path)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-4000:]


def test_documented_symbols_exist_in_the_header():
    """Every bb200_* identifier INTEGRATION.md mentions is declared in include/beluga_b200.h (or is one of its types/macros)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    header = open(os.path.join(ROOT, "include", "beluga_b200.h")).read()
    names = set(re.findall(r"\b(?:bb200|BB200)_[A-Za-z0-9_]+\b", text))
    missing = sorted(n for n in names if not n.endswith("_") and not re.search(r"\b" + re.escape(n) + r"\b", header))
    wildcards = {n for n in missing if any(h.startswith(n) for h in re.findall(r"\bbb200_[a-z0-9_]+", header))}  # `bb200_filter_enqueue_*`
    assert not (set(missing) - wildcards), sorted(set(missing) - wildcards)


def test_documents_point_at_files_that_exist():
    """Paths quoted in the documents (`profiles/...`, `tests/...`, `tools/...`, `include/...`, `oracle/...`, `beluga_b200/...`)."""
    docs = ["DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md"), os.path.join("profiles", "scaling", "README.md"),
            os.path.join("tools", "README.md")]
    missing = []
    for doc in docs:
        text = open(os.path.join(ROOT, doc)).read()
        base_dirs = {"profiles/README.md": "profiles", "profiles/scaling/README.md": os.path.join("profiles", "scaling"), "tools/README.md": "tools"}
        for quoted in re.findall(r"`([^`\n]+)`", text):
            for token in re.split(r"[\s,;()]+", quoted):
                token = token.split("::")[0].rstrip(".:")
                if not re.match(r"^(profiles|tests|tools|include|oracle|beluga_b200)/[A-Za-z0-9_./{}*<>-]+$", token):
                    continue
                if any(ch in token for ch in "*{}<>"):  # patterns such as `profiles/r02_lfm_*`
                    continue
                if token.endswith("/"):
                    token = token[:-1]
                if token in ("oracle/_ref", "oracle/_build"):  # build outputs (the reference itself cannot be compiled here: DESIGN.md)
                    continue
                if not os.path.exists(os.path.join(ROOT, token)):
                    missing.append((doc, token))
        # bare file names in the profiles tables (relative to the document's own directory)
        rel = base_dirs.get(doc.replace(os.sep, "/"))
        if rel:
            for name in re.findall(r"`([A-Za-z0-9_./-]+\.(?:txt|json|csv|sh|py))`", text):
                if "/" in name and not name.startswith("scaling/"):
                    continue
                if not (os.path.exists(os.path.join(ROOT, rel, name)) or os.path.exists(os.path.join(ROOT, name))
                        or os.path.exists(os.path.join(ROOT, "profiles", name)) or os.path.exists(os.path.join(ROOT, "tools", name))):
                    missing.append((doc, name))
    assert not missing, missing


def test_headline_numbers_in_the_documents_are_the_committed_bench_line():
    """DESIGN.md and README.md quote the N = 1 bench line; the figures must be the ones in profiles/r02_bench_n1_final.json."""
    import json

    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_n1_final.json")).read())
    ms, value, e2e = d["ms_per_step"], d["value"], d["e2e"]["value"]
    reweight, frac = d["kernels_ms"]["reweight_lfm"], d["roofline"]["frac"]
    assert d["n_gpus"] == 1 and d["config"]["particles"] == 1_000_000 and d["steps"] >= 20 and d["warmup"] >= 3
    assert abs(ms * value - 1e3) < 1e-6 and 0.0 < frac < 1.0 and reweight < ms
    assert abs(d["roofline"]["achieved"] / d["roofline"]["peak"] - frac) < 1e-9
    assert d["clocks"]["reasons"] == [] and d["gpu_launches"] == 7 * d["steps"]
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    readme = open(os.path.join(ROOT, "README.md")).read()
    for text in (design, readme):
        assert f"{ms:.3f} ms" in text
        assert f"{value:.0f} steps/s" in text
        assert f"{e2e:.0f}" in text
        assert f"{reweight:.3f} ms" in text
    assert f"{frac:.3f}" in design and f"{frac:.2f}" in readme
