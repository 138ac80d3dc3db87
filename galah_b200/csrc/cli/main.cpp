// galah-b200: the `galah cluster` command line over libgalah_b200.so.
//
// Mirrors the reference's flag surface, defaults, argument rules and output formats for the hot path
// (/root/reference/src/cluster_argument_parsing.rs:1603-1757 flags; :544-716 run_cluster_subcommand;
// :718-775 writers; :1491-1512 parse_percentage; genome input flags of bird_tool_utils as documented
// in docs/tools/cluster.md:40-62), so that a `galah cluster ...` invocation can be replayed on the GPU
// and its cluster file compared byte for byte.  Everything below the argument handling is a call into
// the C ABI (include/galah_b200.h); there is no computation in this file and no CPU fallback.
// Not mirrored (out of the hot path): CheckM / CheckM2 / genome-info quality ordering and
// filtering -- genomes are clustered in the order given, which is what the reference does without quality
// input (cluster_argument_parsing.rs:880-883) --, the fastANI backend, --full-help.
#include <dirent.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <string>
#include <vector>

#include "galah_b200.h"

namespace {

struct Options {
    std::vector<std::string> genomes, references;
    std::string fasta_dir, fasta_ext = "fna", fasta_list, reference_list;
    float ani = 95.f, min_af = 15.f, precluster_ani = 90.f;  // crate::DEFAULT_* (src/lib.rs:78-85)
    std::string precluster_method = "skani", cluster_method = "skani";
    bool small_genomes = false, cluster_contigs = false, small_contigs = false, large_contigs = false, low_memory = false;
    int threads = 0, gpus = 1, device = 0;  // threads: 0 = every host core (they only read and inflate files)
    bool quiet = false, print_config = false;
    std::string out_clusters, out_rep_list, out_rep_dir, out_rep_dir_copy;
};

[[noreturn]] void die(const std::string &msg, int code = 1) {
    fprintf(stderr, "%s\n", msg.c_str());
    exit(code);
}

[[noreturn]] void die_lib(const char *what) {
    // the library carries the reference's panic text where the reference panics
    die(std::string(what) + ": " + galah_b200_last_error());
}

void usage(FILE *f) {
    fputs("galah-b200 cluster: cluster FASTA files by average nucleotide identity on B200 GPUs\n\n"
          "Genome input (one or more of):\n"
          "  -f, --genome-fasta-files <PATH>...   -d, --genome-fasta-directory <DIR>\n"
          "  -x, --genome-fasta-extension <EXT>   [default: fna]    --genome-fasta-list <FILE>\n"
          "Clustering:\n"
          "  --ani <F>                    [default: 95]   --min-aligned-fraction <F>  [default: 15]\n"
          "  --precluster-ani <F>         [default: 90]   --precluster-method <skani|finch> [default: skani]\n"
          "  --cluster-method <skani>     [default: skani]  --small-genomes   --low-memory\n"
          "  --cluster-contigs (with --small-contigs or --large-contigs)\n"
          "  --reference-genomes <PATH>... | --reference-genomes-list <FILE>\n"
          "  -t, --threads <N>            host threads for reading files [default: all cores; galah's default is 1]\n"
          "  --gpus <N>                   devices of this box to use [default: 1]   --device <ID> [default: 0]\n"
          "Output (at least one):\n"
          "  -o, --output-cluster-definition <FILE>      representative<TAB>member lines\n"
          "  --output-representative-list <FILE>\n"
          "  --output-representative-fasta-directory <DIR> | --output-representative-fasta-directory-copy <DIR>\n"
          "  -q, --quiet    -h, --help    --version    --print-config (the effective parameters, then exit)\n", f);
}

float parse_f32(const std::string &flag, const char *v) {
    char *end = nullptr;
    const float x = strtof(v, &end);
    if (!*v || *end) die("error: invalid value '" + std::string(v) + "' for '" + flag + "': invalid float literal", 2);
    return x;
}

// parse_percentage (src/cluster_argument_parsing.rs:1491-1512): 1..=100 is a percentage and is divided by
// 100 IN F32; 0..1 is taken as a fraction; anything else is refused.
float parse_percentage(const std::string &parameter, float percentage) {
    if (percentage >= 1.0f && percentage <= 100.0f) return percentage / 100.0f;
    if (!(percentage >= 0.0f && percentage <= 100.0f)) {
        char buf[256];
        snprintf(buf, sizeof buf, "Invalid percentage specified for --%s: '%g'", parameter.c_str(), percentage);
        die(buf);
    }
    return percentage;
}

std::string before_tab(const std::string &s) { return s.substr(0, s.find('\t')); }

std::vector<std::string> read_lines(const std::string &path, const char *what) {
    std::ifstream in(path);
    if (!in) die(std::string("Failed to read ") + what + " file: " + path);
    std::vector<std::string> out;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.find_first_not_of(" \t") == std::string::npos) continue;
        out.push_back(before_tab(line));
    }
    return out;
}

Options parse(int argc, char **argv) {
    Options o;
    auto is_flag = [](const char *a) { return a[0] == '-' && a[1] != 0; };
    for (int i = 2; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&]() -> const char * {
            if (i + 1 >= argc) die("error: a value is required for '" + a + "' but none was supplied", 2);
            return argv[++i];
        };
        auto values = [&](std::vector<std::string> &into) {
            if (i + 1 >= argc || is_flag(argv[i + 1])) die("error: a value is required for '" + a + "' but none was supplied", 2);
            while (i + 1 < argc && !is_flag(argv[i + 1])) into.push_back(argv[++i]);
        };
        if (a == "-f" || a == "--genome-fasta-files") values(o.genomes);
        else if (a == "-d" || a == "--genome-fasta-directory") o.fasta_dir = value();
        else if (a == "-x" || a == "--genome-fasta-extension") o.fasta_ext = value();
        else if (a == "--genome-fasta-list") o.fasta_list = value();
        else if (a == "--ani") o.ani = parse_f32(a, value());
        else if (a == "--min-aligned-fraction") o.min_af = parse_f32(a, value());
        else if (a == "--precluster-ani") o.precluster_ani = parse_f32(a, value());
        else if (a == "--precluster-method") o.precluster_method = value();
        else if (a == "--cluster-method") o.cluster_method = value();
        else if (a == "--small-genomes") o.small_genomes = true;
        else if (a == "--cluster-contigs") o.cluster_contigs = true;
        else if (a == "--small-contigs") o.small_contigs = true;
        else if (a == "--large-contigs") o.large_contigs = true;
        else if (a == "--low-memory") o.low_memory = true;
        else if (a == "--reference-genomes") values(o.references);
        else if (a == "--reference-genomes-list") o.reference_list = value();
        else if (a == "-t" || a == "--threads") o.threads = atoi(value());
        else if (a == "--gpus") o.gpus = atoi(value());
        else if (a == "--device") o.device = atoi(value());
        else if (a == "-o" || a == "--output-cluster-definition") o.out_clusters = value();
        else if (a == "--output-representative-list") o.out_rep_list = value();
        else if (a == "--output-representative-fasta-directory") o.out_rep_dir = value();
        else if (a == "--output-representative-fasta-directory-copy") o.out_rep_dir_copy = value();
        else if (a == "--print-config") o.print_config = true;
        else if (a == "-q" || a == "--quiet") o.quiet = true;
        else if (a == "-v" || a == "--verbose") o.quiet = false;
        else if (a == "-h" || a == "--help") { usage(stdout); exit(0); }
        else if (a == "--checkm-tab-table" || a == "--checkm2-quality-report" || a == "--genome-info" || a == "--run-checkm2" ||
                 a == "--checkm2-db-path" || a == "--min-completeness" || a == "--max-contamination" || a == "--quality-formula")
            die("error: '" + a + "' (quality ordering / filtering) is outside galah-b200: order and filter the genome list "
                "before the call; genomes are clustered in the order given", 2);
        else die("error: unexpected argument '" + a + "' found", 2);
    }
    // clap rules of the reference (cluster_argument_parsing.rs:1660-1757)
    if (o.precluster_method != "skani" && o.precluster_method != "finch")
        die("error: invalid value '" + o.precluster_method + "' for '--precluster-method <precluster-method>'\n  [possible values: skani, finch]", 2);
    if (o.cluster_method == "fastani") die("error: --cluster-method fastani is not available in galah-b200 (skani only)", 2);
    if (o.cluster_method != "skani")
        die("error: invalid value '" + o.cluster_method + "' for '--cluster-method <cluster-method>'\n  [possible values: skani, fastani]", 2);
    if ((o.small_contigs || o.large_contigs) && !o.cluster_contigs)
        die(std::string("error: the following required arguments were not provided:\n  --cluster-contigs"), 2);
    if (o.small_contigs && o.large_contigs)
        die("error: the argument '--small-contigs' cannot be used with '--large-contigs'", 2);
    const bool refs = !o.references.empty() || !o.reference_list.empty();
    if (o.low_memory && refs) die("error: the argument '--low-memory' cannot be used with '--reference-genomes'", 2);
    if (!o.references.empty() && !o.reference_list.empty())
        die("error: the argument '--reference-genomes' cannot be used with '--reference-genomes-list'", 2);
    if (!o.print_config && o.out_clusters.empty() && o.out_rep_list.empty() && o.out_rep_dir.empty() && o.out_rep_dir_copy.empty())
        die("error: the following required arguments were not provided:\n  --output-cluster-definition <output-cluster-definition>\n"
            "  (or one of --output-representative-list, --output-representative-fasta-directory[-copy])", 2);
    if (o.threads < 0) o.threads = 0;
    if (o.gpus < 1) die("error: --gpus must be at least 1", 2);
    return o;
}

std::vector<std::string> genome_paths(const Options &o) {
    std::vector<std::string> paths;
    for (const auto &g : o.genomes) paths.push_back(before_tab(g));
    if (!o.fasta_list.empty())
        for (const auto &g : read_lines(o.fasta_list, "genome fasta list")) paths.push_back(g);
    if (!o.fasta_dir.empty()) {
        DIR *d = opendir(o.fasta_dir.c_str());
        if (!d) die("Failed to open genome fasta directory: " + o.fasta_dir);
        std::vector<std::string> found;
        const std::string suffix = "." + o.fasta_ext;
        while (const dirent *e = readdir(d)) {
            const std::string name = e->d_name;
            if (name.size() > suffix.size() && name.compare(name.size() - suffix.size(), suffix.size(), suffix) == 0)
                found.push_back(o.fasta_dir + (o.fasta_dir.back() == '/' ? "" : "/") + name);
        }
        closedir(d);
        std::sort(found.begin(), found.end());  // read_dir order is not defined; a fixed order makes runs repeatable
        if (found.empty()) die("Found 0 genomes from the genome-fasta-directory, cannot continue.");
        paths.insert(paths.end(), found.begin(), found.end());
    }
    if (paths.empty()) die("No genome fasta files found (use -f, -d or --genome-fasta-list), cannot continue.");
    return paths;
}

// setup_representative_output_directory (cluster_argument_parsing.rs:777-815)
void prepare_directory(const std::string &dir, const char *argument) {
    struct stat st;
    if (stat(dir.c_str(), &st) == 0) {
        if (!S_ISDIR(st.st_mode)) die(std::string("The ") + argument + " path specified (" + dir + ") exists but is not a directory");
        DIR *d = opendir(dir.c_str());
        if (!d) die("Error opening existing output directory " + dir);
        bool empty = true;
        while (const dirent *e = readdir(d))
            if (strcmp(e->d_name, ".") && strcmp(e->d_name, "..")) empty = false;
        closedir(d);
        if (!empty) die(std::string("The ") + argument + " specified (" + dir + ") exists and is not empty");
        return;
    }
    std::string partial;
    for (size_t p = 0; p <= dir.size(); p++)
        if (p == dir.size() || dir[p] == '/') {
            partial = dir.substr(0, p);
            if (!partial.empty() && mkdir(partial.c_str(), 0777) != 0 && errno != EEXIST)
                die(std::string("Error creating ") + argument + " (" + dir + ")");
        }
}

bool exists(const std::string &p) { struct stat st; return lstat(p.c_str(), &st) == 0; }

// write_cluster_reps_to_directory (cluster_argument_parsing.rs:817-849)
void write_reps_to_directory(const galah_b200_clusters_t &cl, const std::vector<std::string> &names, const std::string &dir, bool copy,
                             bool quiet) {
    bool clashed = false;
    for (size_t c = 0; c < cl.n_clusters; c++) {
        const std::string &rep = names[cl.members[cl.offsets[c]]];
        char *abs = realpath(rep.c_str(), nullptr);
        if (!abs) die("Failed to convert representative path into an absolute path: " + rep);
        const std::string base = rep.substr(rep.find_last_of('/') == std::string::npos ? 0 : rep.find_last_of('/') + 1);
        const std::string stab = dir + "/" + base;
        std::string target = stab;
        for (size_t counter = 0; exists(target);) {
            if (!clashed && !quiet)
                fprintf(stderr, "[WARN] One or more sequence files have the same file name (e.g. ). Renaming clashes by adding .1.fna, .2.fna etc.\n");
            clashed = true;
            target = stab + "." + std::to_string(++counter) + ".fna";
        }
        if (copy) {
            std::ifstream in(abs, std::ios::binary);
            std::ofstream out(target, std::ios::binary);
            out << in.rdbuf();
            if (!in || !out) die("Failed to copy representative genome " + rep);
        } else if (symlink(abs, target.c_str()) != 0) {
            die("Failed to create symbolic link to representative genome " + rep);
        }
        free(abs);
    }
}

}  // namespace

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "--version")) { printf("%s\n", galah_b200_version()); return 0; }
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) { usage(argc < 2 ? stderr : stdout); return argc < 2 ? 2 : 0; }
    if (strcmp(argv[1], "cluster"))
        die(std::string("error: unrecognized subcommand '") + argv[1] + "' (galah-b200 implements `cluster`)", 2);
    const Options o = parse(argc, argv);
    std::vector<std::string> genomes = genome_paths(o);

    // run_cluster_subcommand's own checks, in its order (cluster_argument_parsing.rs:570-673)
    if (o.cluster_contigs && !o.small_contigs && !o.large_contigs)
        die("Error: When --cluster-contigs is used, either --small-contigs or --large-contigs must be specified.\n"
            "Use --small-contigs for contigs < 20kb, --large-contigs for contigs >= 20kb.");
    if (o.cluster_contigs && (!o.out_rep_dir.empty() || !o.out_rep_dir_copy.empty()))
        die("Cannot specify --cluster-contigs with --output-representative-fasta-directory or --output-representative-fasta-directory-copy");
    std::vector<std::string> references;
    for (const auto &r : o.references) references.push_back(r);
    if (!o.reference_list.empty()) references = read_lines(o.reference_list, "reference genomes list");
    if (!references.empty() && o.cluster_contigs)
        die("Error: Reference genome clustering is not currently supported with --cluster-contigs");
    if (!references.empty()) {
        if (!o.quiet) fprintf(stderr, "[INFO] Clustering against %zu reference genomes\n", references.size());
        std::vector<std::string> combined(references);  // references first (cluster_argument_parsing.rs:676-686)
        combined.insert(combined.end(), genomes.begin(), genomes.end());
        genomes.swap(combined);
    }
    const bool small = o.cluster_contigs ? o.small_contigs : o.small_genomes;  // determine_small_genomes_setting (:1760-1782)
    const float ani_frac = parse_percentage("ani", o.ani);
    const float af_frac = parse_percentage("min-aligned-fraction", o.min_af);
    const float pre_frac = parse_percentage("precluster-ani", o.precluster_ani);
    const float ani_pct = ani_frac * 100.0f;  // SkaniClusterer.threshold (:1478-1485)
    const float af_pct = af_frac * 100.0f;    // --min-af as the skani callers form it (src/skani.rs:153, 742)
    const bool skip_clusterer = o.precluster_method == o.cluster_method;
    const float pre_pct = (skip_clusterer ? ani_frac : pre_frac) * 100.0f;  // SkaniPreclusterer.threshold (:1300-1352)

    if (o.print_config) {
        // the values the reference's structs would hold (FinchPreclusterer.min_ani, SkaniPreclusterer.threshold,
        // SkaniClusterer.threshold, --min-af as the skani callers form it), nine significant digits: exact for f32
        printf("genomes\t%zu\nreferences\t%zu\nprecluster_method\t%s\ncluster_method\t%s\nskip_clusterer\t%d\n"
               "finch_min_ani\t%.9g\nskani_precluster_threshold\t%.9g\nani_threshold\t%.9g\nmin_af_percent\t%.9g\n"
               "small_genomes\t%d\ncluster_contigs\t%d\nlow_memory\t%d\n",
               genomes.size(), references.size(), o.precluster_method.c_str(), o.cluster_method.c_str(),
               (int)(skip_clusterer || o.cluster_contigs), (double)pre_frac, (double)pre_pct, (double)ani_pct, (double)af_pct,
               (int)small, (int)o.cluster_contigs, (int)o.low_memory);
        return 0;
    }

    // what the reference refuses inside cluster(): no device is needed to say so
    if (o.cluster_contigs && o.precluster_method == "finch") die("finch does not support contig comparisons.");  // src/clusterer.rs:39-41
    if (o.precluster_method == "finch" && (!references.empty() || o.low_memory)) {
        // the reference panics in FinchPreclusterer (src/finch.rs:15, 40): the session entry carries its text
        std::vector<const char *> pp, rp;
        for (const auto &g : genomes) pp.push_back(g.c_str());
        for (const auto &r : references) rp.push_back(r.c_str());
        galah_b200_session_t *s = nullptr;
        galah_b200_pair_t *hits = nullptr;
        size_t n_hits = 0;
        if (galah_b200_session_create(&s)) die_lib("session");
        if (!references.empty())
            galah_b200_session_finch_distances_with_references(s, pp.data(), pp.size(), rp.data(), rp.size(), &hits, &n_hits);
        else
            galah_b200_session_finch_distances(s, pp.data(), pp.size(), pre_frac, 1000, 21, 1, o.threads, &hits, &n_hits);
        die(galah_b200_last_error());
    }

    // outputs are opened before the heavy work, as the reference does (:694-695)
    FILE *f_clusters = nullptr, *f_reps = nullptr;
    if (!o.out_clusters.empty() && !(f_clusters = fopen(o.out_clusters.c_str(), "w")))
        die("Failed to open output cluster definition file " + o.out_clusters);
    if (!o.out_rep_list.empty() && !(f_reps = fopen(o.out_rep_list.c_str(), "w")))
        die("Failed to open output representative list file " + o.out_rep_list);
    if (!o.out_rep_dir.empty()) prepare_directory(o.out_rep_dir, "output-representative-fasta-directory");
    if (!o.out_rep_dir_copy.empty()) prepare_directory(o.out_rep_dir_copy, "output-representative-fasta-directory-copy");

    if (o.gpus > 1 ? galah_b200_init_devices(o.gpus) : galah_b200_init(o.device)) die_lib("galah-b200: no usable B200 device");

    std::vector<const char *> paths;
    for (const auto &g : genomes) paths.push_back(g.c_str());
    std::vector<const char *> ref_paths;
    for (const auto &r : references) ref_paths.push_back(r.c_str());
    std::vector<std::string> names(genomes);
    if (o.cluster_contigs) {
        char **cn = nullptr;
        size_t n_names = 0;
        if (galah_b200_contig_names(paths.data(), paths.size(), &cn, &n_names)) die_lib("reading contig names");
        names.assign(cn, cn + n_names);
        galah_b200_contig_names_free(cn, n_names);
    }
    if (!o.quiet) {
        fprintf(stderr, "[INFO] Clustering %zu genomes ..\n", genomes.size());
        fprintf(stderr, "[INFO] Preclustering with %s and clustering with %s\n", o.precluster_method.c_str(), o.cluster_method.c_str());
    }

    galah_b200_clusters_t cl;
    galah_b200_cluster_stats_t stats;
    memset(&cl, 0, sizeof cl);
    memset(&stats, 0, sizeof stats);
    const size_t n = paths.size();
    if (o.precluster_method == "finch") {
        const int rc = o.gpus > 1
                           ? galah_b200_cluster_files_multi(paths.data(), n, o.gpus, pre_frac, ani_pct, af_pct, small, o.threads, &cl, &stats)
                           : galah_b200_cluster_files(paths.data(), n, pre_frac, ani_pct, af_pct, small, o.threads, &cl, &stats);
        if (rc) die_lib("cluster");
    } else if (references.empty() && !o.low_memory) {
        const int rc = o.gpus > 1 ? galah_b200_cluster_files_skani_multi(paths.data(), n, o.gpus, pre_pct, ani_pct, af_pct, small,
                                                                         o.cluster_contigs, o.threads, &cl, &stats)
                                  : galah_b200_cluster_files_skani(paths.data(), n, pre_pct, ani_pct, af_pct, small, o.cluster_contigs,
                                                                   o.threads, &cl, &stats);
        if (rc) die_lib("cluster");
    } else {
        // low-memory / reference forms of the skani preclusterer (src/skani.rs:229-377, 502-687), then the engine
        // with skip_clusterer (same method names, src/clusterer.rs:32-36)
        galah_b200_session_t *s = nullptr;
        galah_b200_pair_t *hits = nullptr;
        size_t n_hits = 0;
        if (galah_b200_session_create(&s)) die_lib("session");
        const int rc = !references.empty()
                           ? galah_b200_session_skani_distances_with_references(s, paths.data(), n, ref_paths.data(), ref_paths.size(),
                                                                                pre_pct, af_frac, small, o.threads, &hits, &n_hits)
                           : galah_b200_session_skani_distances(s, paths.data(), n, pre_pct, af_frac, small, 1, o.threads, &hits, &n_hits);
        if (rc) die_lib("precluster");
        if (galah_b200_cluster_from_distances(n, hits, n_hits, 1, ani_pct, nullptr, nullptr, &cl)) die_lib("cluster");
        stats.n_precluster_hits = n_hits;
        galah_b200_free(hits);
        galah_b200_session_free(s);
    }
    if (!o.quiet) {
        fprintf(stderr, "[INFO] Found %u preclusters. The largest contained %u genomes\n", cl.n_preclusters, cl.largest_precluster);
        fprintf(stderr, "[INFO] Found %zu genome clusters\n", cl.n_clusters);
    }

    // write_galah_outputs (cluster_argument_parsing.rs:718-775)
    if (f_clusters) {
        for (size_t c = 0; c < cl.n_clusters; c++)
            for (uint64_t m = cl.offsets[c]; m < cl.offsets[c + 1]; m++)
                fprintf(f_clusters, "%s\t%s\n", names[cl.members[cl.offsets[c]]].c_str(), names[cl.members[m]].c_str());
        if (fclose(f_clusters)) die("Failed to write to output clusters file");
    }
    if (!o.out_rep_dir.empty()) write_reps_to_directory(cl, names, o.out_rep_dir, false, o.quiet);
    if (!o.out_rep_dir_copy.empty()) write_reps_to_directory(cl, names, o.out_rep_dir_copy, true, o.quiet);
    if (f_reps) {
        for (size_t c = 0; c < cl.n_clusters; c++) fprintf(f_reps, "%s\n", names[cl.members[cl.offsets[c]]].c_str());
        if (fclose(f_reps)) die("Failed to write to output representative list file");
    }
    if (!o.quiet) fprintf(stderr, "[INFO] Finished printing genome clusters\n");
    galah_b200_clusters_free(&cl);
    return 0;
}
