"""CPU oracle for femo-b200 -- TEST INFRASTRUCTURE ONLY.

What pins it.  The reference's arithmetic lives in un-vendored third-party packages (fenics-dolfinx 0.5.1, basix 0.5.x,
UFL 2022.2, FFCx 0.5.x, PETSc + MUMPS; /root/reference/README.md:21) that cannot be installed here, and the reference
ships no tests or golden vectors, so there is no dolfinx output to compare with: PARITY AGAINST DOLFINX ITSELF IS
UNPINNED.  Everything of the reference that is Python IS run here and pins this package:

  * the form definitions: tests/test_reference_forms.py executes the example scripts' own pdeRes / outputForm / ... and
    motor_pde.py symbolically (sympy UFL slice) and integrates them exactly; tests/golden/symbolic/*.npz hold those
    values and this package reproduces them to 1e-13 (tests/test_oracle_symbolic.py).  For polynomial integrands
    dolfinx's quadrature is exact too, so these are dolfinx's values up to round-off and numbering;
  * the call conventions and constants: femo's own utils_dolfinx.py and csdl_opt / FEA classes run over recording
    stand-ins (tests/test_lower_face.py, tests/test_upper_face.py): tolerances, iteration limits, lifting signs,
    assign-versus-accumulate, the quirks B3 / B4 / B8;
  * reference-held numbers: thick_ref of the beam example, the closed forms and manufactured solutions of the
    Poisson examples (tests/test_oracle.py).

What stays from memory: dolfinx's internal dof / cell numbering, its quadrature points for NON-polynomial integrands
(the sin data of config 2, measured quadrature error 2.6e-10 at cell diameter 0.3) and PETSc's SNES stol default.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
Nothing under femo_b200/ does.
"""
