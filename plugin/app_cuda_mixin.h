/* ----------------------------------------------------------------------
   SPPARKS app-style plugin: `/cuda` variants of ising, potts, potts/neighonly, potts/pin
   (and of ising/single, potts/neigh).

   Drop these files into the SPPARKS src/ directory (or add them to the build as
   plugin/Makefile does), rebuild, and `app_style potts/cuda 100` runs the rejection-KMC
   colour sweep of an unchanged input script on a B200 through libspparks_cuda
   (include/spparks_cuda.h).  Registration goes through style_app.h exactly like the
   reference's own apps (src/app_potts_neighonly.h:18-42, src/Make.sh:8-45).

   AppCudaMixin<Base,MODEL> derives from the reference app (AppIsing, AppPotts, ...) so that
   site_energy(), stats(), dumps, diagnostics and the `pin` command keep working on the host
   copy of the spins; it overrides
     init_app()           validate geometry, create the device lattice   (src/app_lattice.cpp:318)
     iterate_rejection()  the sweep loop, one kernel launch per set      (src/app_lattice.cpp:632-711)
     input_app()          extra commands cuda_device / cuda_seed
   Host spins (iarray[0]) are current whenever control leaves iterate_rejection() or
   output->compute() runs, as SURVEY §8b requires.
------------------------------------------------------------------------- */

#ifndef SPK_APP_CUDA_MIXIN_H
#define SPK_APP_CUDA_MIXIN_H

#include "mpi.h"
#include "string.h"
#include "math.h"
#include "stdlib.h"
#include "stdint.h"
#include "app_lattice.h"
#include "comm_lattice.h"
#include "domain.h"
#include "lattice.h"
#include "random_mars.h"
#include "random_park.h"
// output.h has no include guard in the reference: input.cpp includes it before style_app.h
#include "timer.h"
#include "error.h"

#include "spparks_cuda.h"
#include "spk_cuda_host.h"

namespace SPPARKS_NS {

template <class Base, int MODEL>
class AppCudaMixin : public Base, public SpkCudaHost {
 public:
  AppCudaMixin(class SPPARKS *spk, int narg, char **arg) : Base(spk,narg,arg)
  {
    // Base parsed its arguments only if style matched its own name
    // (src/app_potts.cpp:47, src/app_potts_neighonly.cpp:36), so do it here
    if constexpr (MODEL == SPK_ISING || MODEL == SPK_ISING_SINGLE) {
      if (narg != 1) this->error->all(FLERR,"Illegal app_style command");
      nspins_cuda = 2;
    } else {
      if (narg != 2) this->error->all(FLERR,"Illegal app_style command");
      this->nspins = atoi(arg[1]);
      if (this->nspins <= 0) this->error->all(FLERR,"Illegal app_style command");
      this->dt_sweep = 1.0/this->nspins;
      nspins_cuda = this->nspins;
    }
    this->allow_kmc = 0;          // the device path is rejection KMC only
    this->allow_masking = 1;      // `sweep ... mask yes`: mask bytes live on the device
    handle = NULL;
    host_sync = 1;
    pinned = NULL;
    pinned_bytes = 0;
    perf_sites = 0;
    perf_seconds = 0.0;
    device = 0;
    seed_set = 0;
    seed = 0;
    sweep_counter = 0;
    warned = 0;
    device_current = 0;
    device_only = 0;
    dev_nloop = 1;
    masks_on = 0;
  }

  virtual ~AppCudaMixin()
  {
    if (pinned) spk_cuda_host_unregister(pinned);
    if (handle) spk_cuda_destroy(handle);
  }

  void input_app(char *command, int narg, char **arg)
  {
    if (strcmp(command,"cuda_device") == 0) {
      if (narg != 1) this->error->all(FLERR,"Illegal cuda_device command");
      device = atoi(arg[0]);
    } else if (strcmp(command,"cuda_seed") == 0) {
      if (narg != 1) this->error->all(FLERR,"Illegal cuda_seed command");
      seed = strtoull(arg[0],NULL,10);
      seed_set = 1;
    } else if (strcmp(command,"cuda_set") == 0) {
      // cuda_set site range lo hi  ==  set site range lo hi (src/set.cpp:582-627, loop all,
      // no region/fraction/if), computed on the device: same master-stream draw, same values
      if (narg != 4 || strcmp(arg[0],"site") != 0 || strcmp(arg[1],"range") != 0)
        this->error->all(FLERR,"Illegal cuda_set command (cuda_set site range lo hi)");
      if (this->sites_exist == 0 && !device_only) this->error->all(FLERR,"Set command before sites exist");
      int lo = atoi(arg[2]), hi = atoi(arg[3]);
      int pseed = static_cast<int> (this->ranmaster->uniform()*2147483647);   // new RandomPark(double)
      if (pseed == 0) pseed = 1;
      if (device_only) {      // the lattice lives on the device: nothing to bring back
        check(spk_cuda_set_site_range(handle,pseed,lo,hi));
        if (this->me == 0) {
          if (this->screen) fprintf(this->screen,"  " TAGINT_FORMAT " settings made for site\n",this->nglobal);
          if (this->logfile) fprintf(this->logfile,"  " TAGINT_FORMAT " settings made for site\n",this->nglobal);
        }
        return;
      }
      spk_cuda_handle *tmp = NULL;
      spk_cuda_desc desc;
      fill_desc(desc,0.0);
      check(spk_cuda_create(&desc,&tmp));
      int rc = spk_cuda_set_site_range(tmp,pseed,lo,hi);
      if (!rc) rc = spk_cuda_download_spins(tmp,this->iarray[0],this->nlocal);
      spk_cuda_destroy(tmp);
      check(rc);
      bigint nbig = this->nlocal, allcount;
      MPI_Allreduce(&nbig,&allcount,1,MPI_SPK_BIGINT,MPI_SUM,this->world);
      if (this->me == 0) {
        if (this->screen) fprintf(this->screen,"  " BIGINT_FORMAT " settings made for site\n",allcount);
        if (this->logfile) fprintf(this->logfile,"  " BIGINT_FORMAT " settings made for site\n",allcount);
      }
    } else if (strcmp(command,"cuda_pin") == 0) {
      // cuda_pin fraction multi nthresh  ==  pin fraction multi nthresh of app_style potts/pin
      // (src/app_potts_pin.cpp:52-65, 242-349) computed on the device: same draw from the master
      // generator, same pinned sites.  multi = 1 and nthresh > 0 depend on the order of the attempts:
      // one rank only (the library says so)
      if (narg != 3) this->error->all(FLERR,"Illegal pin command");
      if (this->sites_exist == 0 && !device_only) this->error->all(FLERR,"Pin command before sites exist");
      double pfrac = atof(arg[0]);
      int pmulti = atoi(arg[1]), pthresh = atoi(arg[2]);
      if (pfrac < 0.0 || pfrac > 1.0 || (pmulti != 0 && pmulti != 1) || pthresh < 0)
        this->error->all(FLERR,"Illegal pin command");
      int pseed = static_cast<int> (this->ranmaster->uniform()*2147483647);   // new RandomPark(double), :248
      if (pseed == 0) pseed = 1;
      if (device_only) {
        int64_t np = 0;
        check(spk_cuda_pin_create(handle,pfrac,pmulti,pthresh,pseed,&np));
        return;
      }
      spk_cuda_handle *tmp = NULL;
      spk_cuda_desc desc;
      fill_desc(desc,0.0);
      check(spk_cuda_create(&desc,&tmp));
      int rc = 0;
      if (this->nprocs > 1) {
        char id[SPK_CUDA_COMM_ID_BYTES];
        if (this->me == 0) rc = spk_cuda_comm_id(id);
        MPI_Bcast(id,SPK_CUDA_COMM_ID_BYTES,MPI_CHAR,0,this->world);
        if (!rc) rc = spk_cuda_comm_init(tmp,id);
      }
      int64_t npinned = 0;
      if (!rc) rc = spk_cuda_upload_spins(tmp,this->iarray[0],this->nlocal);
      if (!rc) rc = spk_cuda_pin_create(tmp,pfrac,pmulti,pthresh,pseed,&npinned);
      if (!rc) rc = spk_cuda_download_spins(tmp,this->iarray[0],this->nlocal);
      spk_cuda_destroy(tmp);
      check(rc);
    } else if (strcmp(command,"cuda_create_sites") == 0) {
      // cuda_create_sites  ==  create_sites box  for a lattice that lives on the device ONLY:
      // no per-site host arrays (id, xyz, neighbor lists: 148 bytes per site and 32-bit indices in
      // the reference, src/create_sites.cpp:198-498), so that the BASELINE lattices (1024^3 and up)
      // can be driven from an input script.  Geometry as CreateSites::structured_lattice
      // (:229-252): fully periodic box, a whole number of lattice cells.  What needs the host
      // copy afterwards is refused by the reference itself ("... before sites exist": set, dump,
      // the host diagnostics); cuda_set, cuda_pin, diag_style energy/cuda and cluster/cuda work.
      if (narg != 0) this->error->all(FLERR,"Illegal cuda_create_sites command");
      Domain *domain = this->domain;
      if (domain->box_exist == 0) this->error->all(FLERR,"Create_sites command before simulation box is defined");
      if (this->sites_exist || device_only) this->error->all(FLERR,"Cannot create sites after sites already exist");
      if (domain->lattice == NULL) this->error->all(FLERR,"Cannot create sites with undefined lattice");
      if (domain->nonperiodic) this->error->all(FLERR,"cuda_create_sites needs a fully periodic box");
      int dim = domain->dimension;
      double xl = domain->lattice->xlattice, yl = domain->lattice->ylattice, zl = domain->lattice->zlattice;
      int nx = static_cast<int> (domain->xprd / xl);
      int ny = dim >= 2 ? static_cast<int> (domain->yprd / yl) : 1;
      int nz = dim == 3 ? static_cast<int> (domain->zprd / zl) : 1;
      if (fabs(nx*xl - domain->xprd) > 0.0001*xl || (dim > 1 && fabs(ny*yl - domain->yprd) > 0.0001*yl) ||
          (dim > 2 && fabs(nz*zl - domain->zprd) > 0.0001*zl))
        this->error->all(FLERR,"Periodic box is not a multiple of lattice spacing");
      domain->nx = nx; domain->ny = ny; domain->nz = nz;
      int lat = domain->lattice->style;
      if (spk_cuda_maxneigh(lat) == 0) this->error->all(FLERR,"app_style */cuda does not support this lattice style");
      this->maxneigh = spk_cuda_maxneigh(lat);
      bigint nbig = (bigint) nx * ny * nz * spk_cuda_nbasis(lat);
      if (nbig > MAXTAGINT) this->error->all(FLERR,"Too many sites for the site ID type (tagint)");
      this->nglobal = (tagint) nbig;     // nlocal stays 0 and sites_exist 0: there are no host sites
      spk_cuda_desc desc;
      fill_desc(desc,this->temperature);
      check(spk_cuda_create(&desc,&handle));
      if (this->nprocs > 1) {
        char id[SPK_CUDA_COMM_ID_BYTES];
        if (this->me == 0) check(spk_cuda_comm_id(id));
        MPI_Bcast(id,SPK_CUDA_COMM_ID_BYTES,MPI_CHAR,0,this->world);
        check(spk_cuda_comm_init(handle,id));
      }
      // like iarray[0] after create_sites: every site 0 until `cuda_set` (the reference's init_app
      // refuses such a lattice, and so does ours)
      device_only = 1;
      device_current = 1;
      if (this->me == 0) {
        if (this->screen) fprintf(this->screen,"Creating sites ...\n  " TAGINT_FORMAT " sites on the device (no host copy)\n  "
                                  TAGINT_FORMAT " sites have %d neighbors\n",this->nglobal,this->nglobal,this->maxneigh);
        if (this->logfile) fprintf(this->logfile,"Creating sites ...\n  " TAGINT_FORMAT " sites on the device (no host copy)\n  "
                                   TAGINT_FORMAT " sites have %d neighbors\n",this->nglobal,this->nglobal,this->maxneigh);
      }
    } else if (strcmp(command,"cuda_host_sync") == 0) {
      // yes (default): iarray[0] is refreshed before every Output::compute, as host-side
      //   diagnostics and dumps expect
      // no: the user asserts that nothing reads the host spins during a run (only /cuda
      //   diagnostics and the stats line are active); the lattice comes back when the run ends
      if (narg != 1) this->error->all(FLERR,"Illegal cuda_host_sync command");
      if (strcmp(arg[0],"yes") == 0) host_sync = 1;
      else if (strcmp(arg[0],"no") == 0) host_sync = 0;
      else this->error->all(FLERR,"Illegal cuda_host_sync command");
    } else Base::input_app(command,narg,arg);
  }

  // per-site arrays may have moved (Memory::grow): the pinned registration follows iarray[0]

  void grow_app()
  {
    Base::grow_app();
    if (pinned && pinned != (void *) this->iarray[0]) {
      spk_cuda_host_unregister(pinned);
      pinned = NULL;
    }
  }

  // device lattice (cuda_create_sites): AppLattice::init / setup (src/app_lattice.cpp:133-442) without
  // the per-site machinery (sets, comm lists, masks, site seeds) the host copy would need

  void init()
  {
    if (!device_only) { Base::init(); return; }
    enum{NOSWEEP,RANDOM,RASTER,COLOR,COLOR_STRICT};
    if (this->solve) this->error->all(FLERR,"KMC events are not implemented in app");
    if (this->sweepflag != COLOR)
      this->error->all(FLERR,"app_style */cuda on a device lattice (cuda_create_sites) needs sweep color");
    int dimension = this->domain->dimension;
    this->nsector = 1;
    if (this->sectorflag) {          // src/app_lattice.cpp:161-187
      if (this->nsector_user) this->nsector = this->nsector_user;
      else this->nsector = dimension == 2 ? 4 : 8;
      if (dimension == 3 && ((this->nsector == 2 && (this->domain->procgrid[1] != 1 || this->domain->procgrid[2] != 1)) ||
                             (this->nsector == 4 && this->domain->procgrid[2] != 1)))
        this->error->all(FLERR,"Invalid number of sectors");
      if (dimension == 2 && ((this->nsector == 2 && this->domain->procgrid[1] != 1) || this->nsector == 8))
        this->error->all(FLERR,"Invalid number of sectors");
    }
    this->ncolors = this->domain->lattice->ncolors(this->delevent + this->delpropensity);
    if (this->ncolors == 0) this->error->all(FLERR,"Cannot color this combination of lattice and app");
    this->bothflag = (this->nsector > 1 && this->ncolors > 1) ? 1 : 0;
    if (this->ranapp == NULL) {      // :274-278: the master generator ends up where the host path leaves it
      this->ranapp = new RandomPark(this->ranmaster->uniform());
      double rseed = this->ranmaster->uniform();
      this->ranapp->reset(rseed,this->me,100);
    }
    init_app();
    if (this->dt_sweep == 0.0) this->error->all(FLERR,"App did not set dt_sweep");
    this->output->init(this->time);
  }

  void setup()
  {
    if (!device_only) { Base::setup(); return; }
    this->setup_app();
    int n = 1;                       // :401-410 for sweep color: nloop of every set
    if (this->nstop > 0.0) n = static_cast<int> (this->nstop);
    if (this->tstop > 0.0) n = static_cast<int> (this->tstop/this->dt_sweep);
    dev_nloop = n;
    this->dt_rkmc = (double) n * this->dt_sweep;      // every site is attempted n times per pass: ntotal/nglobal = n
    if (this->dt_rkmc == 0.0) this->error->all(FLERR,"Choice of sector stop led to no rKMC events");
    if (this->stoptime - this->time < this->dt_rkmc) this->dt_rkmc = this->stoptime - this->time;
    this->setup_end_app();
    this->nextoutput = this->output->setup(this->time,this->first_run);
  }

  void init_app()
  {
    Base::init_app();   // spin validity check, dt_sweep of potts/neighonly

    if (device_only) {
      if (!seed_set) {
        seed = (uint64_t) (this->ranmaster->uniform() * 9007199254740992.0);
        seed_set = 1;
      }
      check(spk_cuda_set_seed(handle,seed));
      check(spk_cuda_set_temperature(handle,this->temperature));
      if (!this->Lmask) check(spk_cuda_mask_enable(handle,0));
      else if (!masks_on) check(spk_cuda_mask_enable(handle,1));   // once: the masks live on from run to run (:258-267)
      masks_on = this->Lmask ? 1 : 0;
      // init_app's spin check (src/app_potts.cpp:83-90), where the spins are
      if (spk_cuda_validate_spins(handle) != 0) this->error->one(FLERR,"One or more sites have invalid values");
      return;
    }

    if (!seed_set) {
      // like ranapp (src/app_lattice.cpp:274-278): derived from the `seed` command, once
      seed = (uint64_t) (this->ranmaster->uniform() * 9007199254740992.0);
      seed_set = 1;
    }

    // the reference creates and zeroes its mask array once and keeps it from run to run
    // (src/app_lattice.cpp:258-267): the bytes of the previous run's device lattice are carried over
    unsigned char *saved_mask = NULL;
    if (handle && this->Lmask && masks_on && this->nlocal > 0) {
      saved_mask = new unsigned char[this->nlocal];
      if (spk_cuda_download_mask(handle,saved_mask,this->nlocal)) { delete [] saved_mask; saved_mask = NULL; }
    }
    if (handle) { spk_cuda_destroy(handle); handle = NULL; }

    spk_cuda_desc desc;
    fill_desc(desc,this->temperature);
    check(spk_cuda_create(&desc,&handle));
    if (this->Lmask) {
      check(spk_cuda_mask_enable(handle,1));
      if (saved_mask) check(spk_cuda_upload_mask(handle,saved_mask,this->nlocal));
    }
    delete [] saved_mask;
    masks_on = this->Lmask ? 1 : 0;

    // tri with an odd number of rows: the two basis sites of the middle row fall into different
    // sectors (y = (j+1/2)*sqrt(3) vs ymid, src/app_lattice.cpp:959-974): not a box of cells
    if (desc.lattice == SPK_TRI && this->nsector >= 4 && (slab_hi()-slab_lo()) % 2)
      this->error->all(FLERR,"app_style */cuda: sectors in y need an even number of tri rows per processor");

    // page-lock the host spins once: uploads / downloads then run at full PCIe speed
    if (!pinned && this->nlocal > 0) {
      size_t bytes = (size_t) this->nlocal * sizeof(int);
      if (spk_cuda_host_register(this->iarray[0],bytes) == 0) {
        pinned = this->iarray[0];
        pinned_bytes = bytes;
      }
    }
    device_current = 0;      // the host copy (create_sites, set, pin, a previous run) is the only one
    if (spk_cuda_nlocal(handle) != this->nlocal)
      this->error->one(FLERR,"app_style */cuda: slab size does not match nlocal");

    if (this->nprocs > 1) {
      char id[SPK_CUDA_COMM_ID_BYTES];
      if (this->me == 0) check(spk_cuda_comm_id(id));
      MPI_Bcast(id,SPK_CUDA_COMM_ID_BYTES,MPI_CHAR,0,this->world);
      check(spk_cuda_comm_init(handle,id));
    }
  }

  // replaces AppLattice::iterate_rejection (src/app_lattice.cpp:632-711)

  void iterate_rejection(double stoptime)
  {
    // same enum as src/app_lattice.cpp:40 (file-local there)
    enum{NOSWEEP,RANDOM,RASTER,COLOR,COLOR_STRICT};

    int lat = this->domain->lattice->style;
    int ncolor = spk_cuda_ncolors(lat);
    int strict = (this->sweepflag == COLOR_STRICT);

    if ((this->sweepflag == RANDOM || this->sweepflag == RASTER) && !warned && this->me == 0) {
      this->error->warning(FLERR,"app_style */cuda runs sweep random/raster as colour sweeps "
                           "(same attempts per site and time step; statistically equivalent)");
      warned = 1;
    }

    check(spk_cuda_set_temperature(handle,this->temperature));
    if (!device_only) check(spk_cuda_upload_spins(handle,this->iarray[0],this->nlocal));
    device_current = 1;      // from here on the device is ahead or (after every download) equal
    if (strict) check(spk_cuda_set_siteseeds(handle,this->siteseeds,this->nlocal));
    check(spk_cuda_reset_counters(handle));
    bigint naccept0 = this->naccept;
    bigint nattempt0 = this->nattempt;

    // sets: sectors (or the whole domain) x colours; with RANDOM/RASTER the reference has
    // one set per sector and nloop/nselect chosen so that every site is attempted nstop times

    int nsector = this->nsector;

    // whole-lattice colour sweeps with one pass per set (the common case: `sweep color`,
    // `sector no`) go through spk_cuda_sweeps, which is bit-identical to the per-colour calls
    // below but may fuse the colours of a sweep into one launch

    int whole = (!strict && nsector == 1);
    for (int icolor = 0; whole && icolor < ncolor; icolor++)
      if (set_nloop(0,icolor,ncolor) != 1) whole = 0;

    // the reference's loop sets: sectors (bothflag, or sweep random/raster run as colours), else
    // the colour sets themselves; boundary_clear_mask acts on the loop set (src/app_lattice.cpp:655)
    int sets_are_sectors = !(this->sweepflag == COLOR || this->sweepflag == COLOR_STRICT) || this->bothflag;

    double t_start = MPI_Wtime();
    bigint sweeps_start = this->nsweeps;
    int done = 0;
    while (!done) {
      this->timer->stamp();
      if (whole) check(spk_cuda_sweeps(handle,sweep_counter,1));   // (masks: handled per colour set inside)
      else for (int isector = 0; isector < nsector; isector++) {
        int box[6];
        sector_box(isector,nsector,box);
        if (this->Lmask && sets_are_sectors)
          check(spk_cuda_mask_clear_border(handle,nsector > 1 ? box : NULL,0));
        for (int icolor = 0; icolor < ncolor; icolor++) {
          int nloop = set_nloop(isector,icolor,ncolor);
          if (nloop < 1) continue;
          if (this->nprocs > 1) {
            this->timer->stamp(TIME_SOLVE);
            check(spk_cuda_halo_exchange(handle));      // comm->all(), src/app_lattice.cpp:648-652
            this->timer->stamp(TIME_COMM);
          }
          if (this->Lmask && !sets_are_sectors)
            check(spk_cuda_mask_clear_border(handle,NULL,icolor+1));
          if (strict)
            check(spk_cuda_sweep_color_strict(handle,icolor+1,nsector > 1 ? box : NULL,nloop));
          else
            check(spk_cuda_sweep_color(handle,sweep_counter,icolor+1,nsector > 1 ? box : NULL,nloop));
        }
      }
      sweep_counter++;
      this->nsweeps++;
      this->time += this->dt_rkmc;
      if (this->time >= stoptime) done = 1;

      if (done || this->time >= this->nextoutput) {
        // host copy must be current before host-side diagnostics and dumps read it
        // (cuda_host_sync no: only when the run ends)
        if ((host_sync || done) && !device_only)
          check(spk_cuda_download_spins(handle,this->iarray[0],this->nlocal));
        int64_t acc,att;
        check(spk_cuda_counters(handle,&acc,&att));
        this->naccept = naccept0 + acc;
        this->nattempt = nattempt0 + att;
        this->timer->stamp(TIME_SOLVE);
        this->nextoutput = this->output->compute(this->time,done);
        this->timer->stamp(TIME_OUTPUT);
      }
    }
    if (strict) check(spk_cuda_get_siteseeds(handle,this->siteseeds,this->nlocal));

    // throughput of this run's sweep loop (uploads, downloads and output included), next to
    // the reference's own timing breakdown (src/finish.cpp:30-129)
    perf_seconds += MPI_Wtime() - t_start;
    perf_sites += (double) (this->nsweeps - sweeps_start) * (double) this->nglobal;
    if (this->me == 0 && perf_seconds > 0.0) {
      double gs = perf_sites/perf_seconds*1.0e-9;
      double bytes = 4.0*(ncolor+1);
      char line[256];
      snprintf(line,sizeof line,"CUDA sweep loop: %.4g site-updates in %.4g s = %.4g Gsites/s "
               "(%.4g GB/s of algorithmic lattice traffic at %g bytes/site-update)\n",
               perf_sites,perf_seconds,gs,gs*bytes,bytes);
      if (this->screen) fputs(line,this->screen);
      if (this->logfile) fputs(line,this->logfile);
    }
  }

  // SpkCudaHost: the /cuda diagnostics read the device lattice.  Output::compute is only ever
  // called when host and device agree (after the download above) or before the first sweep of a
  // run (Output::setup), when only the host copy exists.

  spk_cuda_handle *cuda_lattice()
  {
    if (!handle) this->error->one(FLERR,"app_style */cuda: no device lattice before init");
    if (!device_current && !device_only) {
      check(spk_cuda_upload_spins(handle,this->iarray[0],this->nlocal));
      device_current = 1;
    }
    return handle;
  }

 private:
  spk_cuda_handle *handle;
  int device,nspins_cuda,seed_set,warned,device_current,host_sync;
  int device_only;      // the lattice lives on the device only (cuda_create_sites): nlocal = 0 on the host
  int dev_nloop;        // its nloop (setup)
  int masks_on;         // the device lattice carries mask bytes of an earlier run
  void *pinned;
  size_t pinned_bytes;
  double perf_sites,perf_seconds;

  // geometry checks + descriptor of this rank's slab (init_app, cuda_set)

  void fill_desc(spk_cuda_desc &desc, double temperature)
  {
    // the device kernels address sites by lattice index: need a `create_sites box` lattice
    // that is fully periodic (domain->nx,ny,nz set, src/create_sites.cpp:248-252) and
    // one of the five supported styles (enum of src/lattice.cpp:32-33)

    Domain *domain = this->domain;
    if (domain->lattice == NULL || domain->nx == 0)
      this->error->all(FLERR,"app_style */cuda requires create_sites box on a periodic lattice");
    int lat = domain->lattice->style;
    if (spk_cuda_maxneigh(lat) == 0 || this->maxneigh != spk_cuda_maxneigh(lat))
      this->error->all(FLERR,"app_style */cuda does not support this lattice style");

    // one process <-> one GPU; with several ranks every rank must own one slab of the
    // last dimension (processors 1 1 P in 3d, 1 P 1 in 2d)

    int sd = domain->dimension - 1;
    for (int d = 0; d < 3; d++)
      if (d != sd && domain->procgrid[d] != 1)
        this->error->all(FLERR,"app_style */cuda needs a slab decomposition (processors 1 1 P)");
    if (this->nprocs > 1 && domain->myloc[sd] != this->me)
      this->error->all(FLERR,"app_style */cuda needs slab index == MPI rank");

    tagint first = this->nlocal ? this->id[0] : 1;
    for (int i = 0; i < this->nlocal; i++)
      if (this->id[i] != first+i)
        this->error->all(FLERR,"app_style */cuda requires sites ordered by ID");

    desc.model = MODEL;
    desc.lattice = lat;
    desc.nx = domain->nx; desc.ny = domain->ny; desc.nz = domain->nz;
    desc.nspins = nspins_cuda;
    desc.temperature = temperature;
    desc.seed = seed;
    desc.device = device;
    desc.rank = this->me;
    desc.nranks = this->nprocs;
  }

  // this rank's slab of the last dimension in global cells (init_app checked slab index == rank)
  int slab_lo()
  {
    int n = this->domain->dimension == 3 ? this->domain->nz : this->domain->ny;
    return (int) ((bigint) this->me * n / this->nprocs);
  }
  int slab_hi()
  {
    int n = this->domain->dimension == 3 ? this->domain->nz : this->domain->ny;
    return (int) ((bigint) (this->me+1) * n / this->nprocs);
  }
  uint64_t seed;
  uint32_t sweep_counter;

  void check(int rc)
  {
    if (rc) this->error->one(FLERR,spk_cuda_last_error());
  }

  // sector of create_set() (src/app_lattice.cpp:959-984) as a box of global cell indices.  The
  // reference splits at the midpoints of THIS RANK's sub-domain: [0,n) in the undivided
  // dimensions, the rank's own slab in the last one.  x < xmid  <=>  cell index < ceil((lo+hi)/2)
  // for every basis site of the supported lattices (tri rows: init_app refuses odd counts).

  void sector_box(int isector, int nsector, int *box)
  {
    int nx = this->domain->nx, ny = this->domain->ny, nz = this->domain->nz;
    int lo[3] = {0,0,0}, hi[3] = {nx,ny,nz};
    int sd = this->domain->dimension - 1;
    lo[sd] = slab_lo(); hi[sd] = slab_hi();
    for (int d = 0; d < 3; d++) { box[2*d] = lo[d]; box[2*d+1] = hi[d]; }
    if (this->domain->dimension == 2) { box[4] = 0; box[5] = nz; }
    if (nsector == 1) return;
    int mid[3];
    for (int d = 0; d < 3; d++) mid[d] = (lo[d]+hi[d]+1)/2;
    if (isector & 1) box[0] = mid[0]; else box[1] = mid[0];
    if (nsector >= 4) { if (isector & 2) box[2] = mid[1]; else box[3] = mid[1]; }
    if (nsector == 8) { if (isector & 4) box[4] = mid[2]; else box[5] = mid[2]; }
  }

  // nloop of the reference's set for (sector,colour) (src/app_lattice.cpp:401-410)

  int set_nloop(int isector, int icolor, int ncolor)
  {
    enum{NOSWEEP,RANDOM,RASTER,COLOR,COLOR_STRICT};
    if (device_only) return dev_nloop;
    int iset;
    if (this->sweepflag == COLOR || this->sweepflag == COLOR_STRICT) {
      if (this->bothflag) iset = this->nsector + isector*ncolor + icolor;
      else iset = icolor;
      return this->set[iset].nloop;
    }
    // RANDOM / RASTER: one set per sector; nselect/nlocal attempts per site
    iset = isector;
    if (this->sweepflag == RASTER) return this->set[iset].nloop;
    if (this->set[iset].nlocal == 0) return 0;
    // dt_rkmc is computed from nselect (src/app_lattice.cpp:385-417): whole passes only, or the
    // clock and the attempt counters would disagree with the attempts actually made
    if (this->set[iset].nselect % this->set[iset].nlocal)
      this->error->all(FLERR,"app_style */cuda runs sweep random as colour sweeps: "
                       "nstop must give a whole number of attempts per site");
    return this->set[iset].nselect / this->set[iset].nlocal;
  }
};

}

#endif
