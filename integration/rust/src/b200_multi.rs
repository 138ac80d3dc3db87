// UNBUILT: no Rust toolchain exists in the image this was written in (cargo / rustc absent).  The text below is the
// code block of INTEGRATION.md, kept in step with it by tests/test_integration_shim.py, which also checks every
// `galah_b200_*` symbol named here against include/galah_b200.h and the built library.
pub fn cluster_on_all_gpus(genomes: &[&str], pre_ani: f32, ani: f32, min_af: f32, small: bool) -> Vec<Vec<usize>> {
    let n_dev = unsafe { galah_b200_device_count() };
    check(unsafe { galah_b200_init_devices(n_dev) });
    let c_paths = to_c_paths(genomes);
    let mut out = GalahB200Clusters::default();
    check(unsafe { galah_b200_cluster_files_multi(c_paths.as_ptr(), genomes.len(), n_dev, pre_ani, ani * 100.0,
                                                  min_af * 100.0, small as i32, 0, &mut out, std::ptr::null_mut()) });
    take_clusters(out)   // representative first, clusters in the reference's order
}
