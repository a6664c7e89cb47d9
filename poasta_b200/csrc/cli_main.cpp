// `poasta` command line: align / view / stats with the reference's flag set
// (/root/reference/src/bin/poasta.rs:66-161, 276-598), on top of the C ABI (include/poasta_b200.h).
// Graph reuse is -I/--graph as in the reference binary (its README's `-g graph` is the gap-open penalty here too).
// Extensions: --device N, --tiebreak exact|dp.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <unistd.h>
#include <vector>

#include "../../include/poasta_b200.h"

static bool ends_with(const std::string &s, const std::string &suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

static bool read_file(const std::string &path, std::string &out) {
    if (ends_with(path, ".gz")) {  // MultiGzDecoder (bin/poasta.rs:174-188): concatenated members are read through
        gzFile f = gzopen(path.c_str(), "rb");
        if (!f) return false;
        char buf[1 << 16];
        int n;
        while ((n = gzread(f, buf, sizeof buf)) > 0) out.append(buf, n);
        gzclose(f);
        return n == 0;
    }
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    std::ostringstream ss;
    ss << in.rdbuf();
    out = ss.str();
    return true;
}

static bool is_fasta_name(const std::string &p) {
    for (const char *ext : {".fa", ".fa.gz", ".fna", ".fna.gz", ".fasta", ".fasta.gz"})
        if (ends_with(p, ext)) return true;
    return false;
}

static int load_graph_any(const std::string &path, poasta_graph **g) {  // bin/poasta.rs:303-317
    std::string data;
    if (!read_file(path, data)) {
        fprintf(stderr, "Error: could not read %s\n", path.c_str());
        return 1;
    }
    int rc = is_fasta_name(path) ? poasta_graph_from_fasta_msa(data.data(), data.size(), g)
                                 : poasta_graph_load((const uint8_t *)data.data(), data.size(), g);
    if (rc) fprintf(stderr, "Error: could not load graph from %s (status %d)\n", path.c_str(), rc);
    return rc;
}

static int write_output(poasta_graph *g, const std::string &otype, const std::string &opath, bool from_align) {
    char *buf = nullptr;
    size_t len = 0;
    int rc;
    if (otype == "poasta") {
        if (opath.empty() && isatty(STDOUT_FILENO)) {  // bin/poasta.rs:461-467
            fprintf(stderr, "WARNING: not writing binary graph data to terminal standard output!\n");
            return 0;
        }
        rc = poasta_graph_save(g, &buf, &len);
    } else if (otype == "fasta") {
        rc = poasta_graph_to_fasta(g, &buf, &len);
    } else if (otype == "gfa") {
        rc = poasta_graph_to_gfa(g, 0, &buf, &len);
    } else if (otype == "gfav1") {
        rc = poasta_graph_to_gfa(g, 1, &buf, &len);
    } else if (otype == "dot") {
        rc = poasta_graph_to_dot(g, from_align ? 1 : 0, &buf, &len);
    } else {
        fprintf(stderr, "error: invalid value '%s' for '--output-type' [possible values: poasta, dot, fasta, gfa, gfav1]\n", otype.c_str());
        return 2;
    }
    if (rc) {
        fprintf(stderr, "Error: writer failed (status %d)\n", rc);
        return 1;
    }
    if (opath.empty()) {
        fwrite(buf, 1, len, stdout);
    } else {
        std::filesystem::path p(opath);
        if (p.has_parent_path()) std::filesystem::create_directories(p.parent_path());  // bin/poasta.rs:448-452
        std::ofstream out(opath, std::ios::binary);
        if (!out) {
            fprintf(stderr, "Error: cannot create %s\n", opath.c_str());
            poasta_free_buffer(buf);
            return 1;
        }
        out.write(buf, (std::streamsize)len);
    }
    poasta_free_buffer(buf);
    return 0;
}

static std::vector<long> parse_gap_penalties(const std::string &s, bool &ok) {  // bin/poasta.rs:238-243
    std::vector<long> v;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ',')) {
        size_t a = item.find_first_not_of(" \t"), b = item.find_last_not_of(" \t");
        char *end = nullptr;
        std::string t = a == std::string::npos ? "" : item.substr(a, b - a + 1);
        long x = strtol(t.c_str(), &end, 10);
        if (t.empty() || *end || x < 0) ok = false;
        v.push_back(x);
    }
    return v;
}

static void usage() {
    fprintf(stderr,
            "Usage: poasta <COMMAND>\n\nCommands:\n  align  Perform multiple sequence alignment and create or update POA graphs\n"
            "  view   Convert POASTA POA graphs to various output formats\n  stats  Print graph statistics\n\n"
            "poasta align [-I graph] [-o out] [-O poasta|dot|fasta|gfa|gfav1] [-d dir] [-m global|semi-global|ends-free]\n"
            "             [-n 4] [-g 6] [-e 2] [-H dijkstra|mingap|path] [--device 0] [--tiebreak exact|dp] <SEQUENCES>\n"
            "poasta view -O <type> [-o out] <GRAPH>\npoasta stats <GRAPH>\n");
}

int main(int argc, char **argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    size_t i = 0;
    while (i < args.size() && (args[i] == "-v" || args[i] == "--verbose" || args[i] == "-vv")) ++i;  // parsed and unused (bin/poasta.rs:70-71)
    if (i >= args.size()) {
        fprintf(stderr, "Error: No subcommand given.\n");
        usage();
        return 1;
    }
    const std::string cmd = args[i++];
    std::string positional, graph_in, out_path, out_type, span = "global", gopen = "6", gext = "2", heur = "mingap", tiebreak = "exact";
    long mismatch = 4;
    int device = 0;
    auto need = [&](const char *flag) -> std::string {
        if (i >= args.size()) {
            fprintf(stderr, "error: a value is required for '%s'\n", flag);
            exit(2);
        }
        return args[i++];
    };
    while (i < args.size()) {
        const std::string a = args[i++];
        if (a == "-I" || a == "--graph") graph_in = need("--graph");
        else if (a == "-o" || a == "--output") out_path = need("--output");
        else if (a == "-O" || a == "--output-type") out_type = need("--output-type");
        else if (a == "-d" || a == "--debug-output") {
            need("--debug-output");
            fprintf(stderr, "WARNING: --debug-output is accepted but the device path writes no debug dumps\n");
        } else if (a == "-m" || a == "--alignment-span") span = need("--alignment-span");
        else if (a == "-n") mismatch = atol(need("-n").c_str());
        else if (a == "-g") gopen = need("-g");
        else if (a == "-e") gext = need("-e");
        else if (a == "-H" || a == "--heuristic") heur = need("--heuristic");
        else if (a == "--device") device = atoi(need("--device").c_str());
        else if (a == "--tiebreak") tiebreak = need("--tiebreak");
        else if (a == "-v" || a == "--verbose") {
        } else if (a == "-h" || a == "--help") {
            usage();
            return 0;
        } else if (!a.empty() && a[0] == '-' && a.size() > 1) {
            fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str());
            return 2;
        } else positional = a;
    }
    if (positional.empty()) {
        fprintf(stderr, "error: the following required arguments were not provided:\n  <%s>\n", cmd == "align" ? "SEQUENCES" : "GRAPH");
        return 2;
    }
    if (cmd == "view") {
        if (out_type.empty()) {
            fprintf(stderr, "error: the following required arguments were not provided:\n  --output-type <OUTPUT_TYPE>\n");
            return 2;
        }
        poasta_graph *g = nullptr;
        if (load_graph_any(positional, &g)) return 1;
        int rc = write_output(g, out_type, out_path, false);
        poasta_graph_free(g);
        return rc;
    }
    if (cmd == "stats") {  // bin/poasta.rs:560-598: everything goes to stderr
        poasta_graph *g = nullptr;
        if (load_graph_any(positional, &g)) return 1;
        poasta_graph_view v;
        poasta_graph_get_view(g, &v);
        fprintf(stderr, "node_count: %u\nnode_count_with_start: %u\nedge_count: %u\n", poasta_graph_node_count(g),
                poasta_graph_node_count_with_start_and_end(g), poasta_graph_edge_count(g));
        double in_sum = v.pred_off[v.n_nodes], out_sum = v.succ_off[v.n_nodes];
        fprintf(stderr, "avg_in_degree: %.2f\navg_out_degree: %.2f\n", in_sum / v.n_nodes, out_sum / v.n_nodes);
        poasta_graph_free(g);
        return 0;
    }
    if (cmd != "align") {
        fprintf(stderr, "error: unrecognized subcommand '%s'\n", cmd.c_str());
        return 2;
    }
    poasta_params prm;
    memset(&prm, 0, sizeof prm);
    if (heur == "dijkstra" || heur == "DIJKSTRA") prm.heuristic = 0;
    else if (heur == "mingap" || heur == "minimumgapcost") prm.heuristic = 1;
    else if (heur == "path" || heur == "pathaware") prm.heuristic = 2;
    else {
        fprintf(stderr, "Error: Invalid heuristic type. Valid options are: dijkstra, mingap, path\n");
        return 1;
    }
    if (span == "global") prm.mode = 0;
    else if (span == "semi-global" || span == "ends-free") prm.mode = 1;  // both map to EndsFree (bin/poasta.rs:287-301)
    else {
        fprintf(stderr, "error: invalid value '%s' for '--alignment-span'\n", span.c_str());
        return 2;
    }
    bool ok = true;
    auto go = parse_gap_penalties(gopen, ok), ge = parse_gap_penalties(gext, ok);
    if (!ok) {
        fprintf(stderr, "Error: Invalid gap penalty value\n");
        return 1;
    }
    if (go.size() == 2 && ge.size() == 2) {
        if ((uint8_t)ge[0] <= (uint8_t)ge[1]) {  // bin/poasta.rs:336-344: falls back to one-piece GapAffine(n, e1, o1)
            fprintf(stderr, "Warning: gap_extend1 (%d) should be greater than gap_extend2 (%d) for two-piece model\n", (int)(uint8_t)ge[0],
                    (int)(uint8_t)ge[1]);
            fprintf(stderr, "Using standard affine gap model instead.\n");
        } else {
            fprintf(stderr, "Error: the two-piece affine model is outside the device path (DESIGN.md section 6)\n");
            return 1;
        }
    } else if (go.size() != 1 || ge.size() != 1) {
        fprintf(stderr, "Error: Standard affine mode requires exactly 1 value for both gap-open and gap-extend (e.g., -g 6 -e 2)\n");
        return 1;
    }
    prm.mismatch = (uint8_t)mismatch;
    prm.gap_open = (uint8_t)go[0];   // `as u8` like the reference (bin/poasta.rs:331-334,415-416)
    prm.gap_extend = (uint8_t)ge[0];
    prm.tiebreak = tiebreak == "dp" ? 1 : 0;
    poasta_graph *g = nullptr;
    if (!graph_in.empty()) {
        if (load_graph_any(graph_in, &g)) return 1;
    } else {
        g = poasta_graph_new();
    }
    std::string fasta;
    if (!read_file(positional, fasta)) {
        fprintf(stderr, "Error: could not read %s\n", positional.c_str());
        return 1;
    }
    const char *fp = fasta.data();
    size_t fl = fasta.size();
    size_t nseq = 0;
    for (size_t k = 0; k < fl; ++k) nseq += fp[k] == '>' && (k == 0 || fp[k - 1] == '\n');
    std::vector<uint32_t> scores(nseq + 1);
    uint64_t cells = 0;
    int rc = poasta_msa_batch_into(device, &g, &fp, &fl, 1, &prm, scores.data(), &cells);
    if (rc) {
        fprintf(stderr, "Error: alignment failed (poasta_status %d)%s\n", rc, rc == POASTA_ERR_CUDA ? ": no usable CUDA device (there is no CPU fallback)" : "");
        poasta_graph_free(g);
        return 1;
    }
    if (out_type.empty()) out_type = "poasta";
    rc = write_output(g, out_type, out_path, true);
    poasta_graph_free(g);
    return rc;
}
