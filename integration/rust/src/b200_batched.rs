// UNBUILT: no Rust toolchain exists in the image this was written in (cargo / rustc absent).  The text below is the
// code block of INTEGRATION.md, kept in step with it by tests/test_integration_shim.py, which also checks every
// `galah_b200_*` symbol named here against include/galah_b200.h and the built library.
type AniBatchFn = unsafe extern "C" fn(ctx: *mut c_void, reps: *const u32, genomes: *const u32, n: usize,
                                      some: *mut u8, ani: *mut f32) -> c_int;
extern "C" {
    fn galah_b200_cluster_from_distances_batched(n_genomes: usize, hits: *const GalahB200Pair, n_hits: usize,
        ani_threshold: c_float, calculate_ani_batch: AniBatchFn, ctx: *mut c_void, max_waves: u32,
        out: *mut GalahB200Clusters, n_waves: *mut u32) -> c_int;
}
unsafe extern "C" fn trampoline<F: FnMut(&[u32], &[u32], &mut [u8], &mut [f32])>(
        ctx: *mut c_void, reps: *const u32, genomes: *const u32, n: usize, some: *mut u8, ani: *mut f32) -> c_int {
    let f = &mut *(ctx as *mut F);   // reps[x] is the QUERY of pair x (calculate_ani(fasta1 = representative, ..))
    f(std::slice::from_raw_parts(reps, n), std::slice::from_raw_parts(genomes, n),
      std::slice::from_raw_parts_mut(some, n), std::slice::from_raw_parts_mut(ani, n));
    0
}
