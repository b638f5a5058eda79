for mb in 3 4 5 6; do echo -n "minblocks $mb: "; VDBRT_LIB=$PWD/openvdb_b200/libvdbrt_mb$mb.so python tests/_prof_c2.py 3 | tail -1; done
