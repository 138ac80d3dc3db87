// UNBUILT: no Rust toolchain exists in the image this was written in (cargo / rustc absent).  The text below is the
// code block of INTEGRATION.md, kept in step with it by tests/test_integration_shim.py, which also checks every
// `galah_b200_*` symbol named here against include/galah_b200.h and the built library.
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_float, c_int};
use std::sync::Arc;

use crate::sorted_pair_genome_distance_cache::SortedPairGenomeDistanceCache;
use crate::{ClusterDistanceFinder, PreclusterDistanceFinder};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct GalahB200Pair { pub i: u32, pub j: u32, pub common: u32, pub total: u32, pub ani: f32 }
#[repr(C)] pub struct GalahB200Session { _private: [u8; 0] }

extern "C" {
    fn galah_b200_init(device: c_int) -> c_int;
    fn galah_b200_last_error() -> *const c_char;
    fn galah_b200_free(p: *mut std::ffi::c_void);
    fn galah_b200_session_create(out: *mut *mut GalahB200Session) -> c_int;
    fn galah_b200_session_free(s: *mut GalahB200Session);
    fn galah_b200_session_set_clusterer(s: *mut GalahB200Session, small_genomes: c_int) -> c_int;
    // FinchPreclusterer, src/finch.rs:4-46
    fn galah_b200_session_finch_distances(s: *mut GalahB200Session, paths: *const *const c_char, n: usize,
        min_ani: c_float, num_kmers: u32, kmer_length: u8, low_memory: c_int, host_threads: c_int,
        out: *mut *mut GalahB200Pair, n_out: *mut usize) -> c_int;
    // SkaniPreclusterer, src/skani.rs:12-74
    fn galah_b200_session_skani_distances(s: *mut GalahB200Session, paths: *const *const c_char, n: usize,
        threshold_pct: c_float, min_aligned_threshold: c_float, small_genomes: c_int, low_memory: c_int,
        host_threads: c_int, out: *mut *mut GalahB200Pair, n_out: *mut usize) -> c_int;
    fn galah_b200_session_skani_distances_contigs(s: *mut GalahB200Session, paths: *const *const c_char, n: usize,
        contig_names: *const *const c_char, n_names: usize, threshold_pct: c_float, min_aligned_threshold: c_float,
        small_genomes: c_int, host_threads: c_int, out: *mut *mut GalahB200Pair, n_out: *mut usize) -> c_int;
    fn galah_b200_session_skani_distances_with_references(s: *mut GalahB200Session,
        combined: *const *const c_char, n: usize, refs: *const *const c_char, n_refs: usize, threshold_pct: c_float,
        min_aligned_threshold: c_float, small_genomes: c_int, host_threads: c_int,
        out: *mut *mut GalahB200Pair, n_out: *mut usize) -> c_int;
    // SkaniClusterer::calculate_ani, src/skani.rs:708-715
    fn galah_b200_session_calculate_ani(s: *mut GalahB200Session, fasta1: *const c_char, fasta2: *const c_char,
        min_aligned_threshold: c_float, small_genomes: c_int, ani: *mut c_float, is_some: *mut c_int) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(galah_b200_last_error()).to_string_lossy().into_owned() }
}

/// One per `cluster()` call; both trait objects hold an `Arc` of it.
pub struct B200Session(*mut GalahB200Session);
unsafe impl Send for B200Session {}
unsafe impl Sync for B200Session {}   // every entry point locks internally; calculate_ani is re-entrant
impl B200Session {
    pub fn new(device: i32) -> Arc<Self> {
        let mut s = std::ptr::null_mut();
        unsafe {
            if galah_b200_init(device) != 0 { panic!("galah_b200_init failed: {}", last_error()); }
            if galah_b200_session_create(&mut s) != 0 { panic!("{}", last_error()); }
        }
        Arc::new(B200Session(s))
    }
}
impl Drop for B200Session { fn drop(&mut self) { unsafe { galah_b200_session_free(self.0) } } }

fn c_strings(v: &[&str]) -> (Vec<CString>, Vec<*const c_char>) {
    let owned: Vec<CString> = v.iter().map(|p| CString::new(*p).unwrap()).collect();
    let ptrs = owned.iter().map(|p| p.as_ptr()).collect();
    (owned, ptrs)
}
fn take(rc: c_int, out: *mut GalahB200Pair, n: usize) -> SortedPairGenomeDistanceCache {
    if rc != 0 { panic!("{}", last_error()); }   // the reference's own panic text where it has one
    let mut cache = SortedPairGenomeDistanceCache::new();
    for p in unsafe { std::slice::from_raw_parts(out, n) } {
        cache.insert((p.i as usize, p.j as usize), Some(p.ani));   // as src/finch.rs:92 / src/skani.rs:205-208
    }
    unsafe { galah_b200_free(out as *mut std::ffi::c_void) };
    cache
}

/// Same fields as `FinchPreclusterer` (src/finch.rs:4-10) + the session.
pub struct B200FinchPreclusterer {
    pub min_ani: f32, pub num_kmers: usize, pub kmer_length: u8, pub low_memory: bool,
    pub session: Arc<B200Session>,
}
impl PreclusterDistanceFinder for B200FinchPreclusterer {
    fn distances(&self, genome_fasta_paths: &[&str]) -> SortedPairGenomeDistanceCache {
        let (_keep, ptrs) = c_strings(genome_fasta_paths);
        let (mut out, mut n) = (std::ptr::null_mut(), 0usize);
        let rc = unsafe { galah_b200_session_finch_distances(self.session.0, ptrs.as_ptr(), ptrs.len(), self.min_ani,
            self.num_kmers as u32, self.kmer_length, self.low_memory as c_int,
            rayon::current_num_threads() as c_int, &mut out, &mut n) };
        take(rc, out, n)   // low_memory -> rc != 0 with the text of src/finch.rs:15
    }
    fn distances_contigs(&self, _g: &[&str], _c: &[&str]) -> SortedPairGenomeDistanceCache {
        SortedPairGenomeDistanceCache::new()   // src/finch.rs:26-33
    }
    fn distances_with_references(&self, _g: &[&str], _r: &[&str]) -> SortedPairGenomeDistanceCache {
        panic!("Reference genome clustering currently only supported with skani preclusterer")   // src/finch.rs:40
    }
    fn method_name(&self) -> &str { "finch" }   // drives skip_clusterer / the contig panic, src/clusterer.rs:32-41
}

/// Same fields as `SkaniPreclusterer` (src/skani.rs:12-18) + the session.
pub struct B200SkaniPreclusterer {
    pub threshold: f32, pub min_aligned_threshold: f32, pub small_genomes: bool, pub threads: u16,
    pub low_memory: bool, pub session: Arc<B200Session>,
}
impl PreclusterDistanceFinder for B200SkaniPreclusterer {
    fn distances(&self, paths: &[&str]) -> SortedPairGenomeDistanceCache {
        let (_k, p) = c_strings(paths);
        let (mut out, mut n) = (std::ptr::null_mut(), 0usize);
        let rc = unsafe { galah_b200_session_skani_distances(self.session.0, p.as_ptr(), p.len(), self.threshold,
            self.min_aligned_threshold, self.small_genomes as c_int, self.low_memory as c_int,
            self.threads as c_int, &mut out, &mut n) };
        take(rc, out, n)
    }
    fn distances_contigs(&self, paths: &[&str], contig_names: &[&str]) -> SortedPairGenomeDistanceCache {
        let (_k, p) = c_strings(paths);
        let (_kn, names) = c_strings(contig_names);
        let (mut out, mut n) = (std::ptr::null_mut(), 0usize);
        let rc = unsafe { galah_b200_session_skani_distances_contigs(self.session.0, p.as_ptr(), p.len(),
            names.as_ptr(), names.len(), self.threshold, self.min_aligned_threshold, self.small_genomes as c_int,
            self.threads as c_int, &mut out, &mut n) };
        take(rc, out, n)
    }
    fn distances_with_references(&self, combined: &[&str], refs: &[&str]) -> SortedPairGenomeDistanceCache {
        let (_k, p) = c_strings(combined);
        let (_kr, r) = c_strings(refs);
        let (mut out, mut n) = (std::ptr::null_mut(), 0usize);
        let rc = unsafe { galah_b200_session_skani_distances_with_references(self.session.0, p.as_ptr(), p.len(),
            r.as_ptr(), r.len(), self.threshold, self.min_aligned_threshold, self.small_genomes as c_int,
            self.threads as c_int, &mut out, &mut n) };
        take(rc, out, n)
    }
    fn method_name(&self) -> &str { "skani" }
}

/// Same fields as `SkaniClusterer` (src/skani.rs:689-693) + the session.
pub struct B200SkaniClusterer {
    pub threshold: f32,             // percentage
    pub min_aligned_threshold: f32, // fraction; the library multiplies by 100 in f32 as src/skani.rs:733 does
    pub small_genomes: bool,
    pub session: Arc<B200Session>,
}
impl ClusterDistanceFinder for B200SkaniClusterer {
    fn initialise(&self) {
        assert!(self.threshold > 1.0);   // src/skani.rs:696-698
        unsafe { galah_b200_session_set_clusterer(self.session.0, self.small_genomes as c_int) };
    }
    fn method_name(&self) -> &str { "skani" }
    fn get_ani_threshold(&self) -> f32 { self.threshold }
    fn calculate_ani(&self, fasta1: &str, fasta2: &str) -> Option<f32> {
        let (a, b) = (CString::new(fasta1).unwrap(), CString::new(fasta2).unwrap());
        let (mut ani, mut some) = (0f32, 0 as c_int);
        let rc = unsafe { galah_b200_session_calculate_ani(self.session.0, a.as_ptr(), b.as_ptr(),
            self.min_aligned_threshold, self.small_genomes as c_int, &mut ani, &mut some) };
        if rc != 0 { panic!("{}", last_error()); }
        Some(ani)   // skani never yields None: 0.0 when it prints no row (src/skani.rs:760)
    }
}
