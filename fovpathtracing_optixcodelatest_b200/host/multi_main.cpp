// multi_main.cpp -- fovpt_multi: the headless demo of main.cpp as a multi-GPU job, one PROCESS per rank (rank k renders on
// GPU k mod #devices).  The reference is single-GPU (main.cpp:328-330 takes device 0); this is the C++ counterpart of
// bench.py --gpus N for applications that embed SampleRenderer:
//   * every rank loads the same OBJ, builds the same BVH and sets the same camera / probe / gaze;
//   * fovpt_multigpu::partition assigns 32-pixel screen tiles to ranks by foveal sample density, and each rank traces only
//     the launch indices that own pixels of its tiles;
//   * rank 0 owns the shared frame (fovpt_multigpu::SharedFrame); the other ranks' resolve kernels store their pixels into it
//     over peer memory, ordered by device-side flags -- no host synchronisation between the processes while frames render;
//   * rank 0 writes the last frame as a PPM.
// The 128-byte frame handle travels from rank 0 to the others through pipes created before the fork.
//
//   fovpt_multi scene.obj width height frames gaze_x gaze_y ranks out.ppm
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "MultiGpu.h"

namespace {

int run_rank(int rank, int world, int handle_fd, const std::vector<int>& publish_fds, char** argv)
{
    try {
        const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), frames = std::atoi(argv[4]);
        const unsigned gx = (unsigned)std::atoi(argv[5]), gy = (unsigned)std::atoi(argv[6]);
        const char* out = argv[8];
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA capable devices found");

        Model* model = loadOBJ(argv[1]);
        float3 lo = make_float3(1e30f, 1e30f, 1e30f), hi = make_float3(-1e30f, -1e30f, -1e30f);
        for (auto m : model->meshes) for (auto& v : m->vertex) {
            lo = make_float3(fminf(lo.x, v.x), fminf(lo.y, v.y), fminf(lo.z, v.z)); hi = make_float3(fmaxf(hi.x, v.x), fmaxf(hi.y, v.y), fmaxf(hi.z, v.z));
        }
        const float3 centre = make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
        const float extent = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
        sutil::Camera camera(make_float3(centre.x, centre.y, centre.z + 1.4f * extent), centre, make_float3(0.f, 1.f, 0.f), 45.0f, w / float(h));
        ProbeData probe; probe.width = 64; probe.height = 32;
        std::vector<Color> texels((size_t)probe.width * probe.height, make_float4(2.5f, 2.5f, 2.5f, 1.0f));
        probe.data = texels.data();
        probe.BuildCDF();
        fovpt_config cfg; fovpt_default_config(&cfg);           // the settings of fovpt_demo (main.cpp), so both write the same image
        cfg.tmin = 1e-4f * extent; cfg.accumulate = 1; cfg.seed_all_passes_with_subframe = 1; cfg.exposure = 1.0f;

        SampleRenderer sample(model, rank % ndev, &cfg);
        sample.resize(make_int2(w, h));
        sample.setCamera(camera);
        sample.setProbe(probe);
        const fovpt_multigpu::TilePartition part = fovpt_multigpu::partition(w, h, make_uint2(gx, gy), cfg, world);
        fovpt_multigpu::apply(sample, part, rank);

        fovpt_peer_handle handle;
        if (rank == 0) {
            fovpt_multigpu::SharedFrame frame = fovpt_multigpu::SharedFrame::create(sample, world);
            handle = frame.handle();
            for (int fd : publish_fds) if (write(fd, &handle, sizeof(handle)) != (ssize_t)sizeof(handle)) throw std::runtime_error("cannot publish the frame handle");
            std::vector<uint32_t> pixels((size_t)w * h);
            for (int f = 0; f < frames; ++f) {
                sample.launchParams.frame.c = make_uint2(gx, gy);
                sample.render();
                frame.present();                                   // the stream waits for every rank's tiles
                sample.downloadPixels(pixels.data());              // the complete frame
                frame.release();
            }
            const fovpt_stats st = sample.stats();
            std::printf("rank 0 of %d: %d frames %dx%d, own share %.1f %% of the samples (%llu paths)\n", world, frames, w, h,
                        100.0 * part.loads[0] / std::max(1e-9, [&] { double s = 0; for (double l : part.loads) s += l; return s; }()), (unsigned long long)st.samples);
            FILE* fp = std::fopen(out, "wb");
            if (!fp) throw std::runtime_error("cannot write the output image");
            std::fprintf(fp, "P6\n%d %d\n255\n", w, h);
            for (int y = h - 1; y >= 0; --y)
                for (int x = 0; x < w; ++x) { const uint32_t p = pixels[(size_t)y * w + x]; std::fputc(p & 255, fp); std::fputc((p >> 8) & 255, fp); std::fputc((p >> 16) & 255, fp); }
            std::fclose(fp);
        } else {
            if (read(handle_fd, &handle, sizeof(handle)) != (ssize_t)sizeof(handle)) throw std::runtime_error("no frame handle from rank 0");
            fovpt_multigpu::SharedFrame frame = fovpt_multigpu::SharedFrame::open(sample, handle, rank);
            for (int f = 0; f < frames; ++f) {
                sample.launchParams.frame.c = make_uint2(gx, gy);
                sample.render();
                frame.present();                                   // arrival flag behind this rank's pixel stores
            }
            if (fovpt_sync(sample.context()) != 0) throw std::runtime_error(fovpt_last_error(sample.context()));
        }
        delete model;
        std::fflush(stdout);                                       // the ranks leave through _exit()
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "FATAL ERROR (rank " << rank << "): " << e.what() << std::endl;
        return 1;
    }
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc < 9) { std::cerr << "usage: " << argv[0] << " scene.obj width height frames gaze_x gaze_y ranks out.ppm\n"; return 2; }
    const int world = std::atoi(argv[7]);
    if (world < 1 || world > 64) { std::cerr << "ranks must be 1..64\n"; return 2; }
    // pipes first, then fork, and only the children touch CUDA (a forked CUDA context is unusable)
    std::vector<int> rd((size_t)world, -1), wr((size_t)world, -1);
    for (int r = 1; r < world; ++r) { int fd[2]; if (pipe(fd) != 0) { perror("pipe"); return 1; } rd[r] = fd[0]; wr[r] = fd[1]; }
    std::vector<pid_t> kids;
    for (int r = 0; r < world; ++r) {
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 1; }
        if (pid == 0) {
            std::vector<int> publish;
            for (int k = 1; k < world; ++k) { if (r == 0) { publish.push_back(wr[k]); close(rd[k]); } else { close(wr[k]); if (k != r) close(rd[k]); } }
            _exit(run_rank(r, world, rd[r], publish, argv));
        }
        kids.push_back(pid);
    }
    for (int k = 1; k < world; ++k) { close(rd[k]); close(wr[k]); }
    int rc = 0;
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = 1; }
    return rc;
}
