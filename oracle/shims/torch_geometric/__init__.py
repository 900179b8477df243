"""Import-only stub for torch_geometric (reference env.yml: PyG 2.2.0, not installed). TEST INFRASTRUCTURE ONLY.
The HiVT graph-attention stages are OUT OF SCOPE (SURVEY §2 #7,#8); the stubs exist so the reference's encoder
module can be *imported* and its SDE classes (FFunc/GFunc/LSDEFunc) used verbatim."""
