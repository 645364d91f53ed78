"""Import paths PyG pickles its conv layers under (checkpoints/cns.pth).  TEST INFRASTRUCTURE ONLY."""
