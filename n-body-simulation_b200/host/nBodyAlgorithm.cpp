#include "nBodyAlgorithm.hpp"

#include "StateFile.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cctype>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <thread>

nBodyAlgorithm::nBodyAlgorithm(double dt_, double t_end_, double vs, std::string &outDir)
    : outputDirectory(outDir), dt(dt_), t_end(t_end_), visualizationStepWidth(vs) {
    nb_config defaults;
    nb_config_default(&defaults);  // G in AU^3 kg^-1 day^-2, computed as the reference does (nBodyAlgorithm.hpp:55-61)
    G = defaults.G;
}

nBodyAlgorithm::~nBodyAlgorithm() {
    if (ctx) nb_destroy(ctx);
}

void nBodyAlgorithm::check(int status, const char *what) {
    if (status == NB_OK) return;
    std::string msg = std::string(what) + ": " + nb_status_string(status);
    if (ctx && nb_last_error(ctx)[0]) msg += std::string(" (") + nb_last_error(ctx) + ")";
    throw std::runtime_error(msg);
}

namespace {
// Single-node rendezvous for the NCCL unique id: rank 0 writes it to a file, the other ranks wait for it.  The file
// name carries the user id and a per-run nonce (the launcher's run id, else the launcher's pid: all ranks of one
// torchrun share it), the file is created exclusively without following links (mode 0600) and its header repeats the
// nonce, so a file left behind by a crashed run or planted by somebody else is never taken for this run's id.
std::string runNonce() {
    if (const char *p = std::getenv("TORCHELASTIC_RUN_ID")) return p;
    return std::to_string((long) getppid());
}
std::string commFilePath() {
    if (const char *p = std::getenv("NBODY_COMM_FILE")) return p;
    const char *port = std::getenv("MASTER_PORT");
    const char *dir = std::getenv("XDG_RUNTIME_DIR");
    std::string nonce = runNonce();
    for (char &c : nonce)
        if (!std::isalnum((unsigned char) c)) c = '_';
    return std::string(dir && *dir ? dir : "/tmp") + "/nbody_b200_nccl_" + std::to_string((long) getuid()) + "_" +
           (port ? port : "0") + "_" + nonce + ".id";
}
struct CommFileRecord {
    char magic[8];
    char nonce[56];
    uint8_t id[NB_COMM_ID_BYTES];
};
void fillHeader(CommFileRecord &r) {
    std::memset(&r, 0, sizeof r);
    std::memcpy(r.magic, "NBCOMM1", 8);
    std::strncpy(r.nonce, runNonce().c_str(), sizeof r.nonce - 1);
}
void writeCommFile(const std::string &path, const uint8_t *id) {
    CommFileRecord r;
    fillHeader(r);
    std::memcpy(r.id, id, NB_COMM_ID_BYTES);
    const std::string tmp = path + ".tmp";
    ::unlink(path.c_str());   // nobody waits yet for THIS run's file: anything here is stale
    ::unlink(tmp.c_str());
    const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
    if (fd < 0) throw std::runtime_error("cannot create " + tmp);
    const bool ok = ::write(fd, &r, sizeof r) == (ssize_t) sizeof r;
    ::close(fd);
    if (!ok || ::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot write " + path);
}
bool readCommFile(const std::string &path, uint8_t *id) {
    const int fd = ::open(path.c_str(), O_RDONLY | O_NOFOLLOW);
    if (fd < 0) return false;
    struct stat st;
    CommFileRecord r, want;
    fillHeader(want);
    const bool ok = ::fstat(fd, &st) == 0 && st.st_uid == getuid() && S_ISREG(st.st_mode) &&
                    ::read(fd, &r, sizeof r) == (ssize_t) sizeof r && std::memcmp(r.magic, want.magic, 8) == 0 &&
                    std::memcmp(r.nonce, want.nonce, sizeof r.nonce) == 0;
    ::close(fd);
    if (ok) std::memcpy(id, r.id, NB_COMM_ID_BYTES);
    return ok;
}
}  // namespace

void nBodyAlgorithm::openDevice(const SimulationData &d) {
    if (!configuration::use_GPUs)
        std::cout << "Note: --use_gpus=false is accepted but there is no CPU back end; running on the GPU." << std::endl;
    nb_config cfg = configuration::toDeviceConfig(G);
    check(nb_create(&cfg, &ctx), "nb_create");
    if (configuration::worldSize > 1) {
        uint8_t id[NB_COMM_ID_BYTES];
        const std::string path = commFilePath();
        if (configuration::rank == 0) {
            check(nb_comm_get_unique_id(id), "nb_comm_get_unique_id");
            writeCommFile(path, id);
        } else {
            for (int tries = 0; !readCommFile(path, id); ++tries) {
                if (tries > 6000) throw std::runtime_error("timed out waiting for " + path);
                std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
        }
        check(nb_comm_init(ctx, id, configuration::worldSize, configuration::rank), "nb_comm_init");
        if (configuration::rank == 0) std::filesystem::remove(path);
    }
    check(nb_set_bodies(ctx, d.mass.size(), d.mass.data(), d.positions_x.data(), d.positions_y.data(),
                        d.positions_z.data(), d.velocities_x.data(), d.velocities_y.data(), d.velocities_z.data()),
          "nb_set_bodies");
    check(nb_enable_timers(ctx, 1), "nb_enable_timers");
    char name[256];
    check(nb_device_name(ctx, name, sizeof name), "nb_device_name");
    std::string device = name;
    timer.setProperties(description, configuration::numberOfBodies, device);
}

void nBodyAlgorithm::computeEnergy(d_type::int_t currentStep) {
    double e[4];
    check(nb_energy(ctx, e), "nb_energy");
    kineticEnergy[currentStep] = e[0];
    potentialEnergy[currentStep] = e[1];
    totalEnergy[currentStep] = e[2];
    virialEquilibrium[currentStep] = e[3];
}

void nBodyAlgorithm::storeAccelerations(d_type::int_t currentStep) {
    std::vector<double> &norms = acceleration[currentStep];
    norms.resize(configuration::numberOfBodies);
    check(nb_get_acceleration_norms(ctx, norms.data()), "nb_get_acceleration_norms");
}

void nBodyAlgorithm::adjustVelocities(const SimulationData &d) {
    const std::size_t n = configuration::numberOfBodies;
    double mass = 0, px = 0, py = 0, pz = 0;
    for (std::size_t i = 0; i < n; ++i) {
        mass += d.mass[i];
        px += d.mass[i] * d.velocities_x[i];
        py += d.mass[i] * d.velocities_y[i];
        pz += d.mass[i] * d.velocities_z[i];
    }
    const double ux = px / mass, uy = py / mass, uz = pz / mass;
    std::vector<double> &ox = velocities_x[0], &oy = velocities_y[0], &oz = velocities_z[0];
    for (std::size_t i = 0; i < n; ++i) {
        ox[i] -= ux;
        oy[i] -= uy;
        oz[i] -= uz;
    }
}

void nBodyAlgorithm::runTimeLoop(const SimulationData &d, const std::function<void()> &forces) {
    const std::size_t n = configuration::numberOfBodies;
    timer.addTimingSequence("Leapfrog Part 1");
    timer.addTimingSequence("Leapfrog Part 2");

    // step 0 of the output: positions as read, velocities with the mean motion removed; the integrator itself starts
    // from the unadjusted velocities already on the device
    positions_x[0] = d.positions_x; positions_y[0] = d.positions_y; positions_z[0] = d.positions_z;
    velocities_x[0] = d.velocities_x; velocities_y[0] = d.velocities_y; velocities_z[0] = d.velocities_z;
    adjustVelocities(d);

    double time = 0.0, timeSinceLastVisualization = 0.0;
    d_type::int_t currentStep = 0;

    forces();
    if (configuration::compute_energy) computeEnergy(currentStep);
    storeAccelerations(currentStep);
    if (isOutputRank()) std::cout << "Finished initial step " << currentStep << std::endl << std::endl;
    streamStep(currentStep);

    time += dt;
    timeSinceLastVisualization += dt;
    currentStep += 1;

    double ms[NB_T_COUNT];
    double completedTime = 0.0;  // simulated time of the last finished step
    bool kickPending = false;  // second half-kick of the previous (non-visualised) step still to be applied
    while (time <= t_end + 0.000001) {
        const bool visualizeCurrentStep = std::abs(timeSinceLastVisualization - visualizationStepWidth) < 0.000001;

        // A run of steps without output is handed to the device as one batch (nb_advance: inner steps replayed from a
        // CUDA graph, the time loop of a small system is launch bound).  The run is found by stepping the two time
        // accumulators exactly as the loop would, so the visualisation decisions are bit for bit the reference's.
        if (!visualizeCurrentStep && !kickPending && batchAlgorithm >= 0) {
            unsigned run = 0;
            double t = time, since = timeSinceLastVisualization;
            while (t <= t_end + 0.000001 && !(std::abs(since - visualizationStepWidth) < 0.000001) && run < 4096) {
                ++run; t += dt; since += dt;
            }
            if (run >= 6) {
                check(nb_advance(ctx, batchAlgorithm, dt, run, ms), "nb_advance");
                // a tree build that failed in the middle of the batch (node pool, coincident bodies) is reported here
                check(nb_synchronize(ctx), "nb_advance (batch)");
                for (unsigned k = 0; k < run; ++k) {
                    recordForceTimers(ms);
                    timer.addTimeToSequence("Leapfrog Part 1", ms[NB_T_LEAPFROG1]);
                    timer.addTimeToSequence("Leapfrog Part 2", ms[NB_T_LEAPFROG2]);
                }
                completedTime = t - dt;
                time = t;
                timeSinceLastVisualization = since;
                continue;
            }
        }

        // kick-drift; when the previous step needed no output its closing half-kick rides along in the same pass
        check(kickPending ? nb_leapfrog_part2_part1(ctx, dt) : nb_leapfrog_part1(ctx, dt), "leapfrog part 1");
        kickPending = false;

        if (visualizeCurrentStep) {
            std::vector<double> &px = positions_x[currentStep], &py = positions_y[currentStep], &pz = positions_z[currentStep];
            px.resize(n); py.resize(n); pz.resize(n);
            check(nb_get_positions(ctx, px.data(), py.data(), pz.data()), "nb_get_positions");
        }

        forces();

        const double next_time = time + dt;
        const bool last_step = !(next_time <= t_end + 0.000001);
        if (visualizeCurrentStep || last_step) {
            check(nb_leapfrog_part2(ctx, dt), "leapfrog part 2");
        } else {
            kickPending = true;
        }
        check(nb_get_timers(ctx, ms), "nb_get_timers");
        timer.addTimeToSequence("Leapfrog Part 1", ms[NB_T_LEAPFROG1]);
        timer.addTimeToSequence("Leapfrog Part 2", kickPending ? 0.0 : ms[NB_T_LEAPFROG2]);

        if (visualizeCurrentStep) {
            if (isOutputRank()) std::cout << "Finished step " << currentStep << std::endl << std::endl;
            storeAccelerations(currentStep);
            std::vector<double> &vx = velocities_x[currentStep], &vy = velocities_y[currentStep], &vz = velocities_z[currentStep];
            vx.resize(n); vy.resize(n); vz.resize(n);
            check(nb_get_velocities(ctx, vx.data(), vy.data(), vz.data()), "nb_get_velocities");
            if (configuration::compute_energy) computeEnergy(currentStep);
            if (checkpointEveryVisualizedStep && !checkpointPath.empty()) {
                const std::vector<double> *pos[3] = {&positions_x[currentStep], &positions_y[currentStep], &positions_z[currentStep]};
                const std::vector<double> *vel[3] = {&vx, &vy, &vz};
                writeCheckpoint(d, d.start_time + time, pos, vel);
            }
            streamStep(currentStep);
            currentStep += 1;
            timeSinceLastVisualization = 0.0;
        }
        completedTime = time;
        time += dt;
        timeSinceLastVisualization += dt;
    }
    check(nb_synchronize(ctx), "nb_synchronize");
    // the last step always closes with its half-kick, so the device holds a complete (x, v) state here
    if (!checkpointPath.empty()) writeCheckpoint(d, d.start_time + completedTime, nullptr, nullptr);
}

void nBodyAlgorithm::writeCheckpoint(const SimulationData &d, double time, const std::vector<double> *pos[3],
                                     const std::vector<double> *vel[3]) {
    const std::size_t n = configuration::numberOfBodies;
    std::vector<double> fetched[6];
    const double *arrays[7] = {d.mass.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (pos && vel) {
        for (int k = 0; k < 3; ++k) {
            arrays[1 + k] = pos[k]->data();
            arrays[4 + k] = vel[k]->data();
        }
    } else {
        for (auto &v : fetched) v.resize(n);
        check(nb_get_positions(ctx, fetched[0].data(), fetched[1].data(), fetched[2].data()), "nb_get_positions");
        check(nb_get_velocities(ctx, fetched[3].data(), fetched[4].data(), fetched[5].data()), "nb_get_velocities");
        for (int k = 0; k < 6; ++k) arrays[1 + k] = fetched[k].data();
    }
    if (!isOutputRank()) return;  // every rank holds the full state; one writes it
    StateFile::writeArrays(checkpointPath, n, arrays, d.names, d.body_classes, time);
}

// ---- output ------------------------------------------------------------------------------------------------------
namespace {

// orbit classes of https://pdssbn.astro.umd.edu/data_other/objclass.shtml plus STA/DWA/PLA/SAT, numbered as the
// reference does (nBodyAlgorithm.cpp:298-342); unknown -> 0
int orbitClassId(const std::string &c) {
    static const char *const names[] = {"AMO", "APO", "ATE", "IEO", "MCA", "IMB", "MBA", "OMB", "CEN",
                                        "TJN", "TNO", "AST", "PAA", "HYA", "STA", "DWA", "PLA", "SAT"};
    for (int i = 0; i < 18; ++i)
        if (c == names[i]) return i + 1;
    return 0;
}

void openArray(std::ofstream &f, const char *type, const char *name, int components) {
    f << "<DataArray type=\"" << type << "\" Name=\"" << name << "\" NumberOfComponents=\"" << components
      << "\" format=\"ascii\">" << '\n';
}
void openField(std::ofstream &f, const char *name) {
    f << "<DataArray type=\"Float64\" Name=\"" << name << "\" NumberOfTuples=\"1\" format=\"ascii\">" << '\n';
}
const char *const kCloseArray = "</DataArray>";

}  // namespace

void nBodyAlgorithm::prepareOutputDirectory() {
    if (!lastOutputPath.empty()) return;
    // <vs_dir>/<ctime with ' ' -> '_'>/ exactly like the reference (nBodyAlgorithm.cpp:132-140)
    std::time_t now = std::time(nullptr);
    std::string stamp = std::ctime(&now);
    std::replace(stamp.begin(), stamp.end(), ' ', '_');
    stamp.pop_back();  // trailing newline
    lastOutputPath = outputDirectory + '/' + stamp + '/';
    std::filesystem::create_directories(lastOutputPath);
}

void nBodyAlgorithm::writeStepFile(d_type::int_t step, const SimulationData &d) {
    if (binaryOutput) return writeStepFileBinary(step, d);
    const bool energy = configuration::compute_energy;
    std::ofstream f(lastOutputPath + "simulation_step" + std::to_string(step) + ".vtp");
    const std::vector<double> &px = positions_x[step], &py = positions_y[step], &pz = positions_z[step];
    const std::vector<double> &vx = velocities_x[step], &vy = velocities_y[step], &vz = velocities_z[step];
    const std::vector<double> &an = acceleration[step];
    const std::size_t n = px.size();

    f << "<?xml version=\"1.0\"?>" << '\n'
      << "<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">" << '\n'
      << "<PolyData>" << '\n'
      << "<Piece NumberOfPoints=\"" << n << "\" NumberOfVerts=\"" << n << "\">" << '\n'
      << "<Points>" << '\n';
    openArray(f, "Float64", "position", 3);
    for (std::size_t j = 0; j < n; ++j) f << px.at(j) << " " << py.at(j) << " " << pz.at(j) << '\n';
    f << kCloseArray << '\n' << "</Points>" << '\n' << "<PointData>" << '\n';

    openArray(f, "Int32", "body_id", 1);
    for (std::size_t j = 0; j < n; ++j) f << j << '\n';
    f << kCloseArray << '\n';

    openArray(f, "Float64", "velocity", 3);
    for (std::size_t j = 0; j < vx.size(); ++j) f << vx.at(j) << " " << vy.at(j) << " " << vz.at(j) << '\n';
    f << kCloseArray << '\n';

    openArray(f, "Float64", "acceleration", 1);
    for (std::size_t j = 0; j < n; ++j) f << an[j] << '\n';
    f << kCloseArray << '\n';

    openArray(f, "Float64", "mass", 1);
    for (double m : d.mass) f << m << '\n';
    f << kCloseArray << '\n';

    // names as space separated ASCII codes terminated by " 0" (nBodyAlgorithm.cpp:344-351)
    // (bodies read from a binary state file may carry no names / classes: empty name, class 0)
    openArray(f, "String", "name", 1);
    for (const std::string &nm : d.names) {
        for (char c : nm) f << (int) c << ' ';
        f << " 0" << '\n';
    }
    for (std::size_t j = d.names.size(); j < n; ++j) f << " 0" << '\n';
    f << kCloseArray << '\n';

    openArray(f, "Int32", "orbit_class", 1);
    for (const std::string &c : d.body_classes) f << orbitClassId(c) << '\n';
    for (std::size_t j = d.body_classes.size(); j < n; ++j) f << 0 << '\n';
    f << kCloseArray << '\n' << "</PointData>" << '\n' << "<Verts>" << '\n';

    f << "<DataArray type=\"Int64\" Name=\"offsets\">" << '\n';
    for (std::size_t j = 1; j <= n; ++j) f << std::to_string(j) << ' ';
    f << '\n' << kCloseArray << '\n';
    f << "<DataArray type=\"Int64\" Name=\"connectivity\">" << '\n';
    for (std::size_t j = 0; j < n; ++j) f << std::to_string(j) << ' ';
    f << '\n' << kCloseArray << '\n' << "</Verts>" << '\n' << "</Piece>" << '\n' << "<FieldData>" << '\n';

    // energies are 0 when --energy is off (nBodyAlgorithm.cpp:253-284)
    openField(f, "kinetic energy");
    if (energy) f << kineticEnergy[step] << '\n'; else f << 0 << '\n';
    f << kCloseArray << '\n';
    openField(f, "potential energy");
    if (energy) f << potentialEnergy[step] << '\n'; else f << 0 << '\n';
    f << kCloseArray << '\n';
    openField(f, "total energy");
    if (energy) f << totalEnergy[step] << '\n'; else f << 0 << '\n';
    f << kCloseArray << '\n';
    openField(f, "virial equilibrium");
    if (energy) f << virialEquilibrium[step] << '\n'; else f << 0 << '\n';
    f << kCloseArray << '\n' << "</FieldData>" << '\n' << "</PolyData>" << '\n' << "</VTKFile>" << '\n';
}

// Same PolyData as writeStepFile, with every per-body array as an "appended raw" block: after the '_' marker each block
// is a UInt64 byte count followed by the little-endian values; `offset` counts bytes from the marker.  The name array is
// omitted (VTK has no binary string arrays); everything else keeps its name, type and component count.
void nBodyAlgorithm::writeStepFileBinary(d_type::int_t step, const SimulationData &d) {
    const bool energy = configuration::compute_energy;
    std::ofstream f(lastOutputPath + "simulation_step" + std::to_string(step) + ".vtp", std::ios::binary);
    const std::vector<double> &px = positions_x[step], &py = positions_y[step], &pz = positions_z[step];
    const std::vector<double> &vx = velocities_x[step], &vy = velocities_y[step], &vz = velocities_z[step];
    const std::vector<double> &an = acceleration[step];
    const std::uint64_t n = px.size();

    std::uint64_t offset = 0;
    auto declare = [&](const char *type, const char *name, int components, std::uint64_t bytes) {
        f << "<DataArray type=\"" << type << "\" Name=\"" << name << "\"";
        if (components) f << " NumberOfComponents=\"" << components << "\"";
        f << " format=\"appended\" offset=\"" << offset << "\"/>" << '\n';
        offset += 8 + bytes;
    };
    f << "<?xml version=\"1.0\"?>" << '\n'
      << "<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">" << '\n'
      << "<PolyData>" << '\n'
      << "<Piece NumberOfPoints=\"" << n << "\" NumberOfVerts=\"" << n << "\">" << '\n'
      << "<Points>" << '\n';
    declare("Float64", "position", 3, 24 * n);
    f << "</Points>" << '\n' << "<PointData>" << '\n';
    declare("Int32", "body_id", 1, 4 * n);
    declare("Float64", "velocity", 3, 24 * n);
    declare("Float64", "acceleration", 1, 8 * n);
    declare("Float64", "mass", 1, 8 * n);
    declare("Int32", "orbit_class", 1, 4 * n);
    f << "</PointData>" << '\n' << "<Verts>" << '\n';
    declare("Int64", "offsets", 0, 8 * n);
    declare("Int64", "connectivity", 0, 8 * n);
    f << "</Verts>" << '\n' << "</Piece>" << '\n' << "<FieldData>" << '\n';
    f.precision(17);
    const std::pair<const char *, double> fields[4] = {
        {"kinetic energy", energy ? kineticEnergy[step] : 0.0}, {"potential energy", energy ? potentialEnergy[step] : 0.0},
        {"total energy", energy ? totalEnergy[step] : 0.0}, {"virial equilibrium", energy ? virialEquilibrium[step] : 0.0}};
    for (const auto &field : fields) {
        openField(f, field.first);
        f << field.second << '\n' << kCloseArray << '\n';
    }
    f << "</FieldData>" << '\n' << "</PolyData>" << '\n' << "<AppendedData encoding=\"raw\">" << '\n' << "_";

    // blocks, in declaration order; large arrays go through a bounded staging buffer
    constexpr std::uint64_t kChunk = 1 << 16;
    std::vector<double> stage(3 * kChunk);
    auto blockHeader = [&](std::uint64_t bytes) { f.write(reinterpret_cast<const char *>(&bytes), 8); };
    auto interleaved = [&](const std::vector<double> &a, const std::vector<double> &b, const std::vector<double> &c) {
        blockHeader(24 * n);
        for (std::uint64_t base = 0; base < n; base += kChunk) {
            const std::uint64_t m = std::min(kChunk, n - base);
            for (std::uint64_t j = 0; j < m; ++j) {
                stage[3 * j] = a[base + j];
                stage[3 * j + 1] = b[base + j];
                stage[3 * j + 2] = c[base + j];
            }
            f.write(reinterpret_cast<const char *>(stage.data()), (std::streamsize) (24 * m));
        }
    };
    auto generated = [&](auto value, auto make) {  // value: element type tag, make(j) -> element
        using T = decltype(value);
        blockHeader(sizeof(T) * n);
        T *buf = reinterpret_cast<T *>(stage.data());
        for (std::uint64_t base = 0; base < n; base += kChunk) {
            const std::uint64_t m = std::min(kChunk, n - base);
            for (std::uint64_t j = 0; j < m; ++j) buf[j] = make(base + j);
            f.write(reinterpret_cast<const char *>(buf), (std::streamsize) (sizeof(T) * m));
        }
    };
    interleaved(px, py, pz);
    generated(std::int32_t(0), [](std::uint64_t j) { return (std::int32_t) j; });
    interleaved(vx, vy, vz);
    blockHeader(8 * n);
    f.write(reinterpret_cast<const char *>(an.data()), (std::streamsize) (8 * n));
    blockHeader(8 * n);
    f.write(reinterpret_cast<const char *>(d.mass.data()), (std::streamsize) (8 * n));
    generated(std::int32_t(0), [&](std::uint64_t j) {
        return (std::int32_t) (j < d.body_classes.size() ? orbitClassId(d.body_classes[j]) : 0);
    });
    generated(std::int64_t(0), [](std::uint64_t j) { return (std::int64_t) (j + 1); });
    generated(std::int64_t(0), [](std::uint64_t j) { return (std::int64_t) j; });
    f << '\n' << "</AppendedData>" << '\n' << "</VTKFile>" << '\n';
}

void nBodyAlgorithm::enableStreaming(const SimulationData &d) { streamData = &d; }

void nBodyAlgorithm::streamStep(d_type::int_t step) {
    if (!streamData || !isOutputRank()) return;
    prepareOutputDirectory();
    writeStepFile(step, *streamData);
    stepsStreamed = step + 1;
    // keep only what lastState.csv needs, release the rest
    lastPos_x.swap(positions_x[step]); lastPos_y.swap(positions_y[step]); lastPos_z.swap(positions_z[step]);
    positions_x.erase(step); positions_y.erase(step); positions_z.erase(step);
    velocities_x.erase(step); velocities_y.erase(step); velocities_z.erase(step);
    acceleration.erase(step);
}

void nBodyAlgorithm::generateParaViewOutput(const SimulationData &d) {
    if (!isOutputRank()) return;
    prepareOutputDirectory();
    const std::string &base = lastOutputPath;

    timer.exportJSON(base + "times.json");
    outputLastState(base + "lastState.csv");

    std::ofstream pvd(base + "/simulation" + ".pvd");
    pvd << "<?xml version=\"1.0\"?>" << '\n'
        << "<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">"
        << '\n' << "<Collection>" << '\n';

    const d_type::int_t steps = stepsStreamed + (d_type::int_t) positions_x.size();
    for (d_type::int_t step = 0; step < steps; ++step) {
        const std::string name = "simulation_step" + std::to_string(step) + ".vtp";
        pvd << "<DataSet timestep=\"" << step << "\" group=\"\" part=\"0\" file=\"" << name << "\"/>" << '\n';
        if (step >= stepsStreamed) writeStepFile(step, d);
    }
    pvd << "</Collection>" << '\n' << "</VTKFile>" << '\n';
}

void nBodyAlgorithm::outputLastState(const std::string &path) {
    std::ofstream csv(path);
    csv << "position_x, position_y, position_z \n";
    const bool streamed = positions_x.empty();  // streaming mode released the maps and kept the last positions
    const d_type::int_t last = streamed ? 0 : (d_type::int_t) positions_x.rbegin()->first;
    const std::vector<double> &px = streamed ? lastPos_x : positions_x[last];
    const std::vector<double> &py = streamed ? lastPos_y : positions_y[last];
    const std::vector<double> &pz = streamed ? lastPos_z : positions_z[last];
    for (std::size_t j = 0; j < px.size(); ++j) csv << px.at(j) << "," << py.at(j) << "," << pz.at(j) << '\n';
}
