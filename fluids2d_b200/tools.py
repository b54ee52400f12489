"""tools (reference: src/fluids2d/tools.py)."""
from .operators import perpgrad


def set_uv_from_omega(model, omega, u, contravariant=False):
    """velocity from a vertex vorticity: psi = A^-1 omega on the device, then
    u = perpgrad(psi)   (tools.py:6-26)"""
    mesh = model.mesh
    psi = omega * 0
    mesh.poisson_vertices.solve(omega, psi)
    perpgrad(mesh, psi, u, contravariant=contravariant)


def run_twin_experiments(model1, model2, hstack=True):
    """integrate two models with model1's time step (tools.py:29-49); plotting
    stays with the host application, so only the stepping loop is kept"""
    while not model1.time.finished:
        model1.set_dt()
        model2.time.dt = model1.time.dt
        model1.step(1)
        model2.step(1)
        model1.progress()
    model1.progress()
