"""Extract the INPUT vectors the reference's own stage tests use into tests/golden/reference_cases.npz (run in the container that has
/root/reference; the GPU box only sees the committed fixture):

* `many_positions()` — the 512 hand-listed particle positions of the scatter / collect / prepare_tmp / step tests
  (rust/crates/gpu/src/test_util.rs:112-630);
* `test_position_gradients_simple()` — the canonical position gradients (test_util.rs:632-676);
* `specific_positions_and_collider_bits()` — positions + collider bits captured from a failing run (test_util.rs:678-2729);
* the torus collider mesh of the collide tests (rust/crates/gpu/src/torus.rs: vertices(), triangles()).

Only numbers are taken (test vectors); no code.  Usage: python tests/tools/extract_reference_vectors.py"""
import os
import re

import numpy as np

REF = "/root/reference/rust/crates/gpu/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "reference_cases.npz")
NUM = r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?"


def body(text, fn_name):
    start = text.index(f"fn {fn_name}(")
    end = text.index("\n}\n", start)
    return text[start:end]


def main():
    tu = open(os.path.join(REF, "test_util.rs")).read()
    many = np.array([[float(a), float(b), float(c)] for a, b, c in re.findall(rf"Vector3::new\(({NUM}),\s*({NUM}),\s*({NUM})\)", body(tu, "many_positions"))], np.float32)
    spec_body = body(tu, "specific_positions_and_collider_bits")
    spec_pos = np.array([[float(a), float(b), float(c)] for a, b, c in re.findall(rf"Vector3::new\(({NUM}),\s*({NUM}),\s*({NUM})\)", spec_body)], np.float32)
    spec_bits = np.array([int(b) for b in re.findall(r"collider_bits:\s*(\d+)", spec_body)], np.uint32)
    grads = np.array([[float(x) for x in re.findall(NUM, m)] for m in re.findall(r"from_row_slice\(&\[(.*?)\]\)", body(tu, "test_position_gradients_simple"), flags=re.S)], np.float32).reshape(-1, 3, 3)
    to = open(os.path.join(REF, "torus.rs")).read()
    verts = np.array([[float(a), float(b), float(c)] for a, b, c in re.findall(rf"Vector3::new\(({NUM}),\s*({NUM}),\s*({NUM})\)", body(to, "vertices"))], np.float32)
    tris = np.array([[int(a), int(b), int(c)] for a, b, c in re.findall(r"a:\s*(\d+),\s*b:\s*(\d+),\s*c:\s*(\d+)", body(to, "triangles"))], np.uint32)
    if tris.size == 0:
        tris = np.array([[int(a), int(b), int(c)] for a, b, c in re.findall(r"\[\s*(\d+),\s*(\d+),\s*(\d+)\s*\]", body(to, "triangles"))], np.uint32)
    assert many.shape == (512, 3) and spec_pos.shape[0] == spec_bits.shape[0] > 0 and verts.shape[0] > 0 and tris.shape[0] > 0 and int(tris.max()) < verts.shape[0], (many.shape, spec_pos.shape, spec_bits.shape, verts.shape, tris.shape)
    np.savez_compressed(OUT, many_positions=many, specific_positions=spec_pos, specific_bits=spec_bits, position_gradients_rows=grads, torus_vertices=verts, torus_triangles=tris)
    print("wrote", OUT, {k: v.shape for k, v in dict(many_positions=many, specific_positions=spec_pos, specific_bits=spec_bits, position_gradients_rows=grads, torus_vertices=verts, torus_triangles=tris).items()})


if __name__ == "__main__":
    main()
