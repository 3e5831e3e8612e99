"""Host-side mirror of the reference's FE setup (meshes, reference elements, FE spaces, params)."""
