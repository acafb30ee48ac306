// m2venc_tb - C++ host that replays the stimulus of the reference testbench
// (SIM/tb_mpeg2encoder.v:142-274) through the C-ABI of include/m2venc.h:
// for every video: read planar yuv444p frames (TB:210-218), push them 4 pixels per "clock" or in
// bulk (TB:224-235), pulse i_sequence_stop (TB:249-251), and write every o_en word to the .m2v file
// while o_sequence_busy (TB:256-265).  Videos run back to back on ONE encoder instance (TB:150).
//
// usage: m2venc_tb [-XL n] [-YL n] [-VL n] [-Q n] [-P n] [-push4] <in.yuv> <width> <height> <out.m2v> [...more quadruples]
// defaults = testbench defaults: XL=7 YL=6 VECTOR_LEVEL=3 Q_LEVEL=2 i_pframes_count=23 (TB:23-24,98-99,106)
#include "../../include/m2venc.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

int main(int argc, char **argv) {
    int XL = 7, YL = 6, VL = 3, Q = 2, P = 23, push4 = 0, i = 1;
    for (; i < argc && argv[i][0] == '-'; i++) {
        std::string a = argv[i];
        if (a == "-push4") { push4 = 1; continue; }
        if (i + 1 >= argc) break;
        int v = atoi(argv[++i]);
        if (a == "-XL") XL = v; else if (a == "-YL") YL = v; else if (a == "-VL") VL = v; else if (a == "-Q") Q = v; else if (a == "-P") P = v;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if ((argc - i) < 4 || (argc - i) % 4) { fprintf(stderr, "usage: %s [opts] in.yuv W H out.m2v [...]\n", argv[0]); return 2; }
    m2v_encoder *e = nullptr;
    int rc = m2v_create(XL, YL, VL, Q, &e);                                       // reset (TB:144-148)
    if (rc) { fprintf(stderr, "*** m2v_create failed (%d): a B200 is required\n", rc); return 1; }
    for (int nv = 1; i < argc; i += 4, nv++) {
        const char *fin = argv[i], *fout = argv[i + 3];
        const int xs = atoi(argv[i + 1]), ys = atoi(argv[i + 2]);
        printf("start to encode video %d (%4dx%4d)\n", nv, xs, ys);                // TB:173
        FILE *fi = fopen(fin, "rb"), *fo = fopen(fout, "wb");
        if (!fi) { printf("*** couldn't open input file\n"); return 1; }           // TB:175-180
        if (!fo) { printf("*** couldn't open output file\n"); return 1; }
        if (xs < 64 || xs > (16 << XL) || xs % 16) { printf("*** xsize=%4d is invalid\n", xs); return 1; }   // TB:189-194
        if (ys < 64 || ys > (16 << YL) || ys % 16) { printf("*** ysize=%4d is invalid\n", ys); return 1; }   // TB:196-201
        int mbw, mbh;
        if ((rc = m2v_begin(e, xs / 16, ys / 16, P, &mbw, &mbh))) { fprintf(stderr, "begin: %s\n", m2v_last_error(e)); return 1; }
        const size_t fsz = (size_t)xs * ys * 3;
        std::vector<uint8_t> frame(fsz), word(1 << 20);
        long f = 0;
        auto sink = [&]() {                                                       // TB:259-264
            for (;;) {
                size_t n = 0; int last = 0;
                if (m2v_drain(e, word.data(), word.size(), &n, &last)) return;
                if (n) fwrite(word.data(), 1, n, fo);
                if (n < word.size() / 32 * 32) return;
            }
        };
        while (fread(frame.data(), 1, fsz, fi) == fsz) {                          // whole frames only (TB:220, $feof)
            if (push4) {
                const uint8_t *Y = frame.data(), *U = Y + (size_t)xs * ys, *V = U + (size_t)xs * ys;
                for (size_t p = 0; p < (size_t)xs * ys; p += 4) rc |= m2v_push4(e, Y + p, U + p, V + p);
            } else rc = m2v_push_frames(e, frame.data(), 1);
            if (rc) { fprintf(stderr, "push: %s\n", m2v_last_error(e)); return 1; }
            sink();
            f++;
        }
        if ((rc = m2v_stop(e))) { fprintf(stderr, "stop: %s\n", m2v_last_error(e)); return 1; }   // TB:249-251
        sink();
        fclose(fi); fclose(fo);
        printf("end of video %d (%ld frames), busy=%d\n", nv, f, m2v_busy(e));     // TB:270
    }
    m2v_destroy(e);
    return 0;
}
