// ray_voxel_cli.cpp -- the `ray_voxel` executable (boundary B-2): same argv contract, messages
// and exit codes as the reference binary built by examplebuildvoxelgridfrommotion.bat:4,15.
// All work happens behind the C ABI (pvp_ray_voxel_main, include/pvp_b200.h).
#include "../../include/pvp_b200.h"

int main(int argc, char **argv) { return pvp_ray_voxel_main(argc, argv); }
