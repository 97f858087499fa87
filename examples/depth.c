/*
 * C client of libflatgfa: what `fgfa -I graph.gfa depth -d`, `fgfa depth` and
 * `fgfa window-depth PATH SIZE` do, through the C ABI of include/flatgfa.h.
 * The reference ships flatgfa-c/example/example.c for its eight accessors; this is the
 * same kind of program for the depth additions.
 *
 *   gcc -I include examples/depth.c -L pollen_b200/lib -lflatgfa \
 *       -Wl,-rpath,$PWD/pollen_b200/lib -o build/depth_example
 *   build/depth_example graph.gfa [path-name window-size]
 */
#include <stdio.h>
#include <stdlib.h>

#include "flatgfa.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s graph.gfa [path-name window-size]\n", argv[0]);
        return 2;
    }
    flatgfa_t g = flatgfa_parse(argv[1]);
    if (!g) {
        fprintf(stderr, "parse failed: %s\n", flatgfa_last_error());
        return 1;
    }
    uint32_t n = flatgfa_get_segment_count(g);
    uint64_t *depth = malloc(sizeof(uint64_t) * (n ? n : 1));
    uint64_t *uniq = malloc(sizeof(uint64_t) * (n ? n : 1));
    int rc = flatgfa_seg_depth(g, depth, uniq);          /* ops::depth::seg_depth_with_uniq */
    if (rc != 0) {
        fprintf(stderr, "depth failed (%d): %s\n", rc, flatgfa_last_error());
        return 1;
    }
    char *text;
    size_t len;
    if (flatgfa_format_seg_depth(g, depth, uniq, &text, &len) == 0) {   /* SegDepth::emit */
        fwrite(text, 1, len, stdout);
        free(text);
    }

    uint32_t n_paths = flatgfa_path_count(g);
    uint64_t *lengths = malloc(sizeof(uint64_t) * (n_paths ? n_paths : 1));
    double *means = malloc(sizeof(double) * (n_paths ? n_paths : 1));
    if (flatgfa_path_depth(g, NULL, n_paths, lengths, means) == 0 &&    /* ops::depth::path_depth */
        flatgfa_format_path_depth(g, NULL, n_paths, lengths, means, &text, &len) == 0) {
        fwrite(text, 1, len, stdout);
        free(text);
    }

    if (argc >= 4) {                                     /* ops::window_depth::window_depth */
        rc = flatgfa_window_depth(g, argv[2], strtoull(argv[3], NULL, 10), &text, &len);
        if (rc != 0) {
            fprintf(stderr, "window depth failed (%d): %s\n", rc, flatgfa_last_error());
            return 1;
        }
        fwrite(text, 1, len, stdout);
        free(text);
    }
    free(depth);
    free(uniq);
    free(lengths);
    free(means);
    flatgfa_free(g);
    return 0;
}
