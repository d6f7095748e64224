/* C (not C++) consumer of include/pcr.h: the header must compile as plain C, and the host-only exports must be
 * callable without a GPU.  Built and run by tests/test_capi_exports.py::test_header_is_plain_c_and_links. */
#include <stdio.h>
#include <string.h>

#include "pcr.h"

#define CHECK(cond)                                             \
    do {                                                        \
        if (!(cond)) {                                          \
            fprintf(stderr, "abi_check: %s failed\n", #cond);   \
            return 1;                                           \
        }                                                       \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    CHECK(pcr_version() >= 100);
    CHECK(sizeof(pcr_hyp_record) == 128);
    pcr_align_params prm;
    pcr_align_default_params(&prm);
    CHECK(prm.voxel_size == 0.3 && prm.ransac_max_iter == 30 && prm.icp_max_iter == 30);
    CHECK(pcr_kernel_class_count() > 5 && pcr_kernel_class_name(0) != NULL);

    /* PLY round trip through the C entry points (argv[1] = scratch path) */
    float xyzw[3][4] = {{1.5f, -2.25f, 3.0f, 0.f}, {0.1f, 0.2f, 0.3f, 0.f}, {-7.f, 8.f, 9.5f, 0.f}};
    float nrm[3][4] = {{0.f, 0.f, 1.f, 0.f}, {0.f, 1.f, 0.f, 0.f}, {1.f, 0.f, 0.f, 0.f}};
    unsigned char rgb[9] = {255, 180, 0, 0, 166, 237, 1, 2, 3};
    char err[128];
    for (int binary = 0; binary < 2; binary++) {
        CHECK(pcr_ply_write(argv[1], &xyzw[0][0], 3, &nrm[0][0], rgb, binary, err, (int)sizeof err) == PCR_OK);
        pcr_ply_info info;
        CHECK(pcr_ply_probe(argv[1], &info, err, (int)sizeof err) == PCR_OK);
        CHECK(info.n_vertex == 3 && info.has_normals == 1 && info.has_colors == 1 && info.format == binary);
        float back[3][4], nback[3][4];
        double x64[3][3];
        memset(back, 0xff, sizeof back);
        CHECK(pcr_ply_read(argv[1], 3, &back[0][0], &nback[0][0], &x64[0][0], 0, NULL, err, (int)sizeof err) == PCR_OK);
        CHECK(memcmp(back, xyzw, sizeof back) == 0 && memcmp(nback, nrm, sizeof nback) == 0);
        CHECK(x64[2][1] == 8.0 && (float)x64[1][0] == 0.1f);
        CHECK(pcr_ply_read(argv[1], 2, &back[0][0], NULL, NULL, 0, NULL, err, (int)sizeof err) == PCR_ERR_INVALID);
    }
    CHECK(pcr_ply_probe("/nonexistent/dir/x.ply", NULL, err, (int)sizeof err) == PCR_ERR_INVALID); /* null info */
    pcr_ply_info info;
    CHECK(pcr_ply_probe("/nonexistent/dir/x.ply", &info, err, (int)sizeof err) == PCR_ERR_IO && strlen(err) > 0);

    /* the RANSAC replay is host-only: an empty wave consumes its hypotheses and changes nothing else */
    pcr_reg_result st;
    memset(&st, 0, sizeof st);
    st.transformation[0] = st.transformation[5] = st.transformation[10] = st.transformation[15] = 1.0;
    st.best_hyp = -1;
    st.est_k = 1000;
    int stop = -1;
    pcr_hyp_record none;
    memset(&none, 0, sizeof none);
    CHECK(pcr_ransac_scan(&none, 0, 0, 256, 50, 100, 0.999, pcr_ransac_k_d(0.0075, 100), &st, &stop) == PCR_OK);
    CHECK(st.hyp_evaluated == 256 && st.best_hyp == -1 && stop == 0);
    /* context-taking calls refuse a null context instead of crashing */
    CHECK(pcr_align_files(NULL, argv[1], argv[1], &prm, NULL) == PCR_ERR_INVALID);
    CHECK(pcr_destroy(NULL) == PCR_OK);
    printf("abi_check ok\n");
    return 0;
}
