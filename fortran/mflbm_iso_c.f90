!=======================================================================================================
! mflbm_iso_c.f90 -- ISO_C_BINDING layer between the unchanged MF-LBM Fortran driver and libmflbm.so
!
! This is the reference-side binding of include/mflbm.h (the C ABI of the B200-native time-step hot
! path).  It is NOT compiled in this repository's image (no Fortran compiler is installed there); it is
! the file a maintainer adds to multiphase_3D/0.src/ (and, with the g*/phi members dropped, to
! singlephase_3D/0.src/) -- see INTEGRATION.md.  All logic lives on the C side; everything below is
! mechanical marshalling of module variables (MP/Module.F90) into the C structs.
!
! Part 1: module mflbm_c      -- interfaces, one per export of include/mflbm.h
! Part 2: module mflbm_glue   -- context handle + helpers that fill mflbm_config / mflbm_arrays from
!                                Misc_module / Fluid_singlephase / Fluid_multiphase / mpi_variable
! Part 3: drop-in subroutine bodies replacing the reference kernels' callers:
!           main_iteration_kernel   (MP/Main_multiphase.F90:341-486)
!           color_gradient          (MP/Phase_gradient.F90:5-204)
!           compute_macro_vars      (MP/Misc.F90:372-430)
!           device part of monitor  (MP/Monitor.F90:27-107), cal_saturation (MP/Monitor.F90:512-550),
!           monitor_breakthrough    (MP/Monitor.F90:472-507)
!=======================================================================================================
module mflbm_c
    use, intrinsic :: iso_c_binding
    implicit none

    integer(c_int), parameter :: MFLBM_OK = 0
    integer(c_int), parameter :: MFLBM_SOLVER_SINGLEPHASE = 0, MFLBM_SOLVER_MULTIPHASE = 1

    ! struct mflbm_config (include/mflbm.h) -- member order and types must match exactly
    type, bind(c) :: mflbm_config
        integer(c_int32_t) :: struct_size
        integer(c_int32_t) :: solver
        integer(c_int32_t) :: nx, ny, nz
        integer(c_int32_t) :: nxGlobal, nyGlobal, nzGlobal
        integer(c_int32_t) :: idz, npz
        integer(c_int32_t) :: jper, kper
        integer(c_int32_t) :: domain_wall_status_z_min, domain_wall_status_z_max
        integer(c_int32_t) :: inlet_BC, outlet_BC
        integer(c_int32_t) :: porous_plate_cmd, Z_porous_plate
        integer(c_int32_t) :: mrt
        integer(c_int32_t) :: iz_async
        integer(c_int32_t) :: num_solid_boundary, num_fluid_boundary
        integer(c_int32_t) :: device
        integer(c_int32_t) :: use_nccl
        integer(c_int32_t) :: kernel_variant
        integer(c_int32_t) :: reserved_i(7)
        real(c_double) :: la_nui1, la_nui2
        real(c_double) :: gamma, beta, force_Z, phi_inlet, sa_inject, relaxation, uin_avg, rho_in, rho_out
        real(c_double) :: s_e, s_e2, s_q, s_nu, s_pi, s_t
        real(c_double) :: reserved_d(8)
        integer(c_signed_char) :: nccl_unique_id(128)
    end type mflbm_config

    ! struct mflbm_arrays: host-array bundle; c_null_ptr members are skipped
    type, bind(c) :: mflbm_arrays
        type(c_ptr) :: f(19)
        type(c_ptr) :: g(19)
        type(c_ptr) :: phi, phi_old
        type(c_ptr) :: cn_x, cn_y, cn_z, c_norm
        type(c_ptr) :: curv
        type(c_ptr) :: u, v, w, rho
        type(c_ptr) :: walls
        type(c_ptr) :: w_in
        type(c_ptr) :: f_convec_bc, g_convec_bc, phi_convec_bc
        type(c_ptr) :: solid_boundary_nodes, fluid_boundary_nodes
    end type mflbm_arrays

    ! struct mflbm_geometry_config (device geometry preprocessing, SURVEY 8(f) item 1)
    type, bind(c) :: mflbm_geometry_config
        integer(c_int32_t) :: struct_size
        integer(c_int32_t) :: nxGlobal, nyGlobal, nzGlobal
        integer(c_int32_t) :: wk0, wk1
        integer(c_int32_t) :: idz, npz
        integer(c_int32_t) :: iper, jper, kper
        integer(c_int32_t) :: device
        real(c_double) :: theta
    end type mflbm_geometry_config

    interface
        integer(c_int) function mflbm_geometry_preprocess(cfg, walls_window, solid, num_solid, fluid, num_fluid, &
                                                          num_solid_scanned, num_fluid_scanned) &
            bind(c, name="mflbm_geometry_preprocess")
            import :: c_int, c_ptr, c_int32_t, c_int64_t, mflbm_geometry_config
            type(mflbm_geometry_config), intent(in) :: cfg
            type(c_ptr), value :: walls_window              ! c_loc(walls_global(1,1,wk0)), integer(kind=1)
            type(c_ptr), intent(out) :: solid, fluid        ! malloc'ed lists, release with mflbm_geometry_free
            integer(c_int32_t), intent(out) :: num_solid, num_fluid
            integer(c_int64_t), intent(out) :: num_solid_scanned, num_fluid_scanned
        end function
        subroutine mflbm_geometry_free(list) bind(c, name="mflbm_geometry_free")
            import :: c_ptr
            type(c_ptr), value :: list
        end subroutine
        integer(c_int) function mflbm_output_begin(ctx, what) bind(c, name="mflbm_output_begin")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: what      ! 1 = phi, 2 = u,v,w,rho (after compute_macro_vars), 3 = both
        end function
        integer(c_int) function mflbm_output_end(ctx, host) bind(c, name="mflbm_output_end")
            import :: c_int, c_ptr, mflbm_arrays
            type(c_ptr), value :: ctx
            type(mflbm_arrays), intent(in) :: host
        end function
        integer(c_int) function mflbm_checkpoint_begin(ctx) bind(c, name="mflbm_checkpoint_begin")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx          ! returns 0: device snapshot taken, the step loop may go on; 1: context frozen
        end function
        integer(c_int) function mflbm_checkpoint_fetch(ctx, host) bind(c, name="mflbm_checkpoint_fetch")
            import :: c_int, c_ptr, mflbm_arrays
            type(c_ptr), value :: ctx
            type(mflbm_arrays), intent(in) :: host
        end function
        integer(c_int) function mflbm_checkpoint_end(ctx) bind(c, name="mflbm_checkpoint_end")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function mflbm_create(cfg, ctx) bind(c, name="mflbm_create")
            import :: c_int, c_ptr, mflbm_config
            type(mflbm_config), intent(in) :: cfg
            type(c_ptr), intent(out) :: ctx
        end function
        subroutine mflbm_destroy(ctx) bind(c, name="mflbm_destroy")
            import :: c_ptr
            type(c_ptr), value :: ctx
        end subroutine
        type(c_ptr) function mflbm_last_error(ctx) bind(c, name="mflbm_last_error")
            import :: c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function mflbm_upload(ctx, host) bind(c, name="mflbm_upload")
            import :: c_int, c_ptr, mflbm_arrays
            type(c_ptr), value :: ctx
            type(mflbm_arrays), intent(in) :: host
        end function
        integer(c_int) function mflbm_download(ctx, host) bind(c, name="mflbm_download")
            import :: c_int, c_ptr, mflbm_arrays
            type(c_ptr), value :: ctx
            type(mflbm_arrays), intent(in) :: host
        end function
        integer(c_int) function mflbm_step(ctx, ntime) bind(c, name="mflbm_step")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: ntime
        end function
        integer(c_int) function mflbm_run(ctx, ntime0, nsteps) bind(c, name="mflbm_run")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), value :: ntime0, nsteps
        end function
        integer(c_int) function mflbm_color_gradient(ctx) bind(c, name="mflbm_color_gradient")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function mflbm_compute_macro_vars(ctx) bind(c, name="mflbm_compute_macro_vars")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function mflbm_monitor(ctx, tk, tk_len) bind(c, name="mflbm_monitor")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: tk(*)
            integer(c_int), value :: tk_len
        end function
        integer(c_int) function mflbm_cal_saturation(ctx, v1, v2) bind(c, name="mflbm_cal_saturation")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: v1, v2
        end function
        ! streamed step: host inlet profile in, saturation sums of the PREVIOUS streamed step out (include/mflbm.h)
        integer(c_int) function mflbm_step_streamed(ctx, ntime, w_in_host, v1, v2, have_prev) bind(c, name="mflbm_step_streamed")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: ntime
            type(c_ptr), value :: w_in_host   ! c_loc(w_in) of a pinned array, or c_null_ptr
            real(c_double), intent(out) :: v1, v2
            integer(c_int), intent(out) :: have_prev
        end function
        integer(c_int) function mflbm_stream_flush(ctx, v1, v2) bind(c, name="mflbm_stream_flush")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: v1, v2
        end function
        integer(c_int) function mflbm_chain_info(ctx, fused, reject_mask) bind(c, name="mflbm_chain_info")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
            integer(c_int), intent(out) :: fused, reject_mask
        end function
        integer(c_int) function mflbm_monitor_breakthrough(ctx, cnt) bind(c, name="mflbm_monitor_breakthrough")
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value :: ctx
            integer(c_int32_t), intent(out) :: cnt
        end function
        integer(c_int) function mflbm_monitor_steady_phasefield(ctx, umax_sq, d_phi_max) &
                bind(c, name="mflbm_monitor_steady_phasefield")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: umax_sq, d_phi_max
        end function
        integer(c_int) function mflbm_monitor_steady_capillarypressure(ctx, umax_sq, pre_w, pre_nw, i_w, i_nw) &
                bind(c, name="mflbm_monitor_steady_capillarypressure")
            import :: c_int, c_ptr, c_double, c_int32_t
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: umax_sq, pre_w, pre_nw
            integer(c_int32_t), intent(out) :: i_w, i_nw
        end function
        integer(c_int) function mflbm_set_parameter(ctx, name, val) bind(c, name="mflbm_set_parameter")
            import :: c_int, c_ptr, c_char, c_double
            type(c_ptr), value :: ctx
            character(kind=c_char), intent(in) :: name(*)
            real(c_double), value :: val
        end function
        integer(c_int) function mflbm_sync(ctx) bind(c, name="mflbm_sync")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function mflbm_timer_start(ctx) bind(c, name="mflbm_timer_start")
            import :: c_int, c_ptr
            type(c_ptr), value :: ctx
        end function
        integer(c_int) function mflbm_timer_stop(ctx, elapsed_ms) bind(c, name="mflbm_timer_stop")
            import :: c_int, c_ptr, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: elapsed_ms
        end function
        integer(c_int) function mflbm_nccl_unique_id(id) bind(c, name="mflbm_nccl_unique_id")
            import :: c_int, c_signed_char
            integer(c_signed_char), intent(out) :: id(128)
        end function
    end interface
end module mflbm_c

!=======================================================================================================
module mflbm_glue
    use, intrinsic :: iso_c_binding
    use mflbm_c
    implicit none
    type(c_ptr), save :: mflbm_handle = c_null_ptr
contains

    ! the reference's error convention: MPI_Barrier + mpi_abort (MP/IO_multiphase.F90:543-545)
    subroutine mflbm_check(rc, where)
        use mpi_variable
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: where
        integer :: ierr
        if (rc /= MFLBM_OK) then
            write(*,*) 'mflbm error ', rc, ' in ', where
            call MPI_Barrier(MPI_COMM_WORLD, ierr)
            call mpi_abort(MPI_COMM_WORLD, 1, ierr)
        endif
    end subroutine

    ! replaces "!$acc data copy(...) copyin(...)" (MP/Main_multiphase.F90:104-115) and setDevice (:70)
    subroutine mflbm_enter_data(device_num)
        use Misc_module
        use Fluid_singlephase
        use Fluid_multiphase
        use mpi_variable
        integer, intent(in) :: device_num
        type(mflbm_config) :: cfg
        type(mflbm_arrays) :: h
        integer :: ierr
        cfg%struct_size = int(c_sizeof(cfg), c_int32_t)
        cfg%solver = MFLBM_SOLVER_MULTIPHASE
        cfg%nx = nx; cfg%ny = ny; cfg%nz = nz
        cfg%nxGlobal = nxGlobal; cfg%nyGlobal = nyGlobal; cfg%nzGlobal = nzGlobal
        cfg%idz = idz; cfg%npz = npz
        cfg%jper = jper; cfg%kper = kper
        cfg%domain_wall_status_z_min = domain_wall_status_z_min
        cfg%domain_wall_status_z_max = domain_wall_status_z_max
        cfg%inlet_BC = inlet_BC; cfg%outlet_BC = outlet_BC
        cfg%porous_plate_cmd = porous_plate_cmd; cfg%Z_porous_plate = Z_porous_plate
        cfg%mrt = 2                                   ! MP/preprocessor.h "#define mrt 2"
        cfg%iz_async = iz_async
        cfg%num_solid_boundary = num_solid_boundary; cfg%num_fluid_boundary = num_fluid_boundary
        cfg%device = device_num
        cfg%use_nccl = merge(1, 0, npz > 1)
        cfg%kernel_variant = 0
        cfg%reserved_i = 0
        cfg%la_nui1 = la_nui1; cfg%la_nui2 = la_nui2
        cfg%gamma = gamma; cfg%beta = beta; cfg%force_Z = force_Z; cfg%phi_inlet = phi_inlet
        cfg%sa_inject = sa_inject; cfg%relaxation = relaxation; cfg%uin_avg = uin_avg
        cfg%rho_in = rho_in; cfg%rho_out = rho_out
        cfg%s_e = 0d0; cfg%s_e2 = 0d0; cfg%s_q = 0d0; cfg%s_nu = 0d0; cfg%s_pi = 0d0; cfg%s_t = 0d0
        cfg%reserved_d = 0d0
        cfg%nccl_unique_id = 0
        if (npz > 1) then                             ! replaces MPI_CART_CREATE for the z ring (MP/Mpi_misc.F90:19-38)
            if (id == 0) call mflbm_check(mflbm_nccl_unique_id(cfg%nccl_unique_id), 'mflbm_nccl_unique_id')
            call MPI_Bcast(cfg%nccl_unique_id, 128, MPI_BYTE, 0, MPI_COMM_VGRID, ierr)
        endif
        call mflbm_check(mflbm_create(cfg, mflbm_handle), 'mflbm_create')
        call mflbm_fill_arrays(h, .true.)
        call mflbm_check(mflbm_upload(mflbm_handle, h), 'mflbm_upload')
    end subroutine

    ! replaces "!$acc end data" (MP/Main_multiphase.F90:325)
    subroutine mflbm_exit_data()
        call mflbm_destroy(mflbm_handle)
        mflbm_handle = c_null_ptr
    end subroutine

    ! c_loc of the module arrays (the arrays need the TARGET attribute in MP/Module.F90)
    subroutine mflbm_fill_arrays(h, with_geometry)
        use Misc_module
        use Fluid_singlephase
        use Fluid_multiphase
        type(mflbm_arrays), intent(out) :: h
        logical, intent(in) :: with_geometry
        h%f(1)  = c_loc(f0);  h%f(2)  = c_loc(f1);  h%f(3)  = c_loc(f2);  h%f(4)  = c_loc(f3);  h%f(5)  = c_loc(f4)
        h%f(6)  = c_loc(f5);  h%f(7)  = c_loc(f6);  h%f(8)  = c_loc(f7);  h%f(9)  = c_loc(f8);  h%f(10) = c_loc(f9)
        h%f(11) = c_loc(f10); h%f(12) = c_loc(f11); h%f(13) = c_loc(f12); h%f(14) = c_loc(f13); h%f(15) = c_loc(f14)
        h%f(16) = c_loc(f15); h%f(17) = c_loc(f16); h%f(18) = c_loc(f17); h%f(19) = c_loc(f18)
        h%g(1)  = c_loc(g0);  h%g(2)  = c_loc(g1);  h%g(3)  = c_loc(g2);  h%g(4)  = c_loc(g3);  h%g(5)  = c_loc(g4)
        h%g(6)  = c_loc(g5);  h%g(7)  = c_loc(g6);  h%g(8)  = c_loc(g7);  h%g(9)  = c_loc(g8);  h%g(10) = c_loc(g9)
        h%g(11) = c_loc(g10); h%g(12) = c_loc(g11); h%g(13) = c_loc(g12); h%g(14) = c_loc(g13); h%g(15) = c_loc(g14)
        h%g(16) = c_loc(g15); h%g(17) = c_loc(g16); h%g(18) = c_loc(g17); h%g(19) = c_loc(g18)
        h%phi = c_loc(phi); h%phi_old = c_null_ptr
        ! steady_state_option 2: phi_old (= phi after initialization_new_multi, MP/Init_multiphase.F90:341-347) goes up with
        ! the rest, so that the first monitor_multiphase_steady_phasefield measures the change since the start like the
        ! reference; without it the library seeds phi_old from phi at the first monitor call
        if (with_geometry .and. steady_state_option == 2) h%phi_old = c_loc(phi_old)
        h%cn_x = c_null_ptr; h%cn_y = c_null_ptr; h%cn_z = c_null_ptr; h%c_norm = c_null_ptr; h%curv = c_null_ptr
        h%u = c_null_ptr; h%v = c_null_ptr; h%w = c_null_ptr; h%rho = c_null_ptr
        h%walls = c_null_ptr; h%w_in = c_null_ptr
        h%solid_boundary_nodes = c_null_ptr; h%fluid_boundary_nodes = c_null_ptr
        h%f_convec_bc = c_null_ptr; h%g_convec_bc = c_null_ptr; h%phi_convec_bc = c_null_ptr
        if (outlet_BC == 1) then
            h%f_convec_bc = c_loc(f_convec_bc); h%g_convec_bc = c_loc(g_convec_bc); h%phi_convec_bc = c_loc(phi_convec_bc)
        endif
        if (with_geometry) then                       ! copyin(walls,w_in,solid_boundary_nodes,fluid_boundary_nodes)
            h%walls = c_loc(walls); h%w_in = c_loc(w_in)
            if (num_solid_boundary > 0) h%solid_boundary_nodes = c_loc(solid_boundary_nodes)
            if (num_fluid_boundary > 0) h%fluid_boundary_nodes = c_loc(fluid_boundary_nodes)
        endif
    end subroutine

    ! replaces "!$acc update host(f0..f18,g0..g18,phi)" of save_checkpoint (MP/IO_multiphase.F90:572-575)
    subroutine mflbm_update_host_checkpoint()
        type(mflbm_arrays) :: h
        call mflbm_fill_arrays(h, .false.)
        call mflbm_check(mflbm_download(mflbm_handle, h), 'mflbm_download(checkpoint)')
    end subroutine

    ! Staged variant of the same update (MP/IO_multiphase.F90:572-575): call mflbm_checkpoint_stage() at the step whose
    ! state is to be saved (it returns at once when the device snapshot fits; the step loop goes on), and
    ! mflbm_checkpoint_collect() right before save_checkpoint writes the arrays: the device-to-host copies run on a copy
    ! stream and overlap the steps queued in between.  ntime+1 in the file header is the step of the stage call.
    subroutine mflbm_checkpoint_stage()
        integer(c_int) :: rc
        rc = mflbm_checkpoint_begin(mflbm_handle)
        if (rc < 0) call mflbm_check(rc, 'mflbm_checkpoint_begin')
    end subroutine
    subroutine mflbm_checkpoint_collect()
        type(mflbm_arrays) :: h
        call mflbm_fill_arrays(h, .false.)
        call mflbm_check(mflbm_checkpoint_fetch(mflbm_handle, h), 'mflbm_checkpoint_fetch')
        call mflbm_check(mflbm_checkpoint_end(mflbm_handle), 'mflbm_checkpoint_end')
    end subroutine

    ! replaces "!$acc update host(u,v,w,phi,rho)" after compute_macro_vars (MP/IO_multiphase.F90:686,792,857)
    subroutine mflbm_update_host_macro()
        use Fluid_singlephase
        use Fluid_multiphase
        type(mflbm_arrays) :: h
        integer :: q
        do q = 1, 19
            h%f(q) = c_null_ptr; h%g(q) = c_null_ptr
        enddo
        h%phi = c_loc(phi); h%phi_old = c_null_ptr
        h%cn_x = c_null_ptr; h%cn_y = c_null_ptr; h%cn_z = c_null_ptr; h%c_norm = c_null_ptr; h%curv = c_null_ptr
        h%u = c_loc(u); h%v = c_loc(v); h%w = c_loc(w); h%rho = c_loc(rho)
        h%walls = c_null_ptr; h%w_in = c_null_ptr
        h%f_convec_bc = c_null_ptr; h%g_convec_bc = c_null_ptr; h%phi_convec_bc = c_null_ptr
        h%solid_boundary_nodes = c_null_ptr; h%fluid_boundary_nodes = c_null_ptr
        call mflbm_check(mflbm_download(mflbm_handle, h), 'mflbm_download(macro)')
    end subroutine
end module mflbm_glue

!=======================================================================================================
! Part 3: drop-in bodies.  Each replaces the body of the reference subroutine of the same name; the
! callers (program main_multiphase, benchmark, monitor tail, IO routines) stay as they are.
!=======================================================================================================

! MP/Main_multiphase.F90:341-486 -- kernels, halo exchange, BCs and color_gradient for this ntime parity
subroutine main_iteration_kernel
    use Misc_module, only: ntime
    use mflbm_glue
    implicit none
    call mflbm_check(mflbm_step(mflbm_handle, int(ntime, c_int)), 'mflbm_step')
end subroutine main_iteration_kernel

! MP/Phase_gradient.F90:5-204 (only the pre-loop call at MP/Main_multiphase.F90:120 still reaches this)
subroutine color_gradient
    use mflbm_glue
    implicit none
    call mflbm_check(mflbm_color_gradient(mflbm_handle), 'mflbm_color_gradient')
end subroutine color_gradient

! MP/Misc.F90:372-430 -- u,v,w,rho stay on the device; IO routines fetch them with mflbm_update_host_macro
subroutine compute_macro_vars
    use mflbm_glue
    implicit none
    call mflbm_check(mflbm_compute_macro_vars(mflbm_handle), 'mflbm_compute_macro_vars')
end subroutine compute_macro_vars

! Device part of monitor (MP/Monitor.F90:27-107): fills tk(1:7*nz+3) exactly as the reference packs it at
! :92-106 (fl1,fl2,vol1,vol2,mass1,mass2,pre per z plane, then umax,usq1,usq2).  The reference's rank-0
! accumulation, saturation / Ca / pressure-drop arithmetic and file output (:108-277) follow unchanged,
! reading fl1(k)=tk(k), fl2(k)=tk(nz+k), ... instead of "!$acc update host".
subroutine monitor_device_part(umax, usq1, usq2)
    use Misc_module
    use Fluid_multiphase
    use mpi_variable
    use mflbm_glue
    implicit none
    real(kind=8), intent(out) :: umax, usq1, usq2
    integer :: k
    call mflbm_check(mflbm_monitor(mflbm_handle, tk, int(7*nz+3, c_int)), 'mflbm_monitor')
    do k = 1, nz
        fl1(k) = tk(k);        fl2(k) = tk(nz+k);     vol1(k) = tk(2*nz+k); vol2(k) = tk(3*nz+k)
        mass1(k) = tk(4*nz+k); mass2(k) = tk(5*nz+k); pre(k) = tk(6*nz+k)
    enddo
    umax = tk(7*nz+1); usq1 = tk(7*nz+2); usq2 = tk(7*nz+3)
end subroutine monitor_device_part

! Device part of cal_saturation (MP/Monitor.F90:527-538): slab sums v1,v2; the MPI_REDUCE and the division
! (:540-548) follow unchanged.
subroutine cal_saturation_device_part(v1, v2)
    use mflbm_glue
    implicit none
    real(kind=8), intent(out) :: v1, v2
    call mflbm_check(mflbm_cal_saturation(mflbm_handle, v1, v2), 'mflbm_cal_saturation')
end subroutine cal_saturation_device_part

! Device part of monitor_breakthrough (MP/Monitor.F90:483-495): integer count on plane nz-1 of the last slab
subroutine monitor_breakthrough_device_part(outlet_phase1_sum)
    use mflbm_glue
    implicit none
    integer, intent(out) :: outlet_phase1_sum
    integer(c_int32_t) :: cnt
    call mflbm_check(mflbm_monitor_breakthrough(mflbm_handle, cnt), 'mflbm_monitor_breakthrough')
    outlet_phase1_sum = cnt
end subroutine monitor_breakthrough_device_part

! geometry_preprocessing_new (MP/Geometry_preprocessing.F90:9-512), called from MP/Main_multiphase.F90:98 on every rank
! (the reference does the work on rank 0 and broadcasts): the body below replaces it.  Each rank hands the library the
! planes of walls_global around its slab and receives its LOCAL lists, which it copies into the allocatable module
! arrays solid_boundary_nodes / fluid_boundary_nodes (MP/Module.F90:96,104) exactly as :424-507 does.
subroutine geometry_preprocessing_new
    use, intrinsic :: iso_c_binding
    use mflbm_c
    use mflbm_glue
    use Misc_module
    use Fluid_multiphase
    use mpi_variable
    implicit none
    type(mflbm_geometry_config) :: gc
    type(c_ptr) :: ps, pf
    type(indirect_solid_boundary_nodes), pointer :: s(:)
    type(indirect_fluid_boundary_nodes), pointer :: f(:)
    integer(c_int32_t) :: ns, nf
    integer(c_int64_t) :: gs, gf
    gc%struct_size = int(c_sizeof(gc), c_int32_t)
    gc%nxGlobal = nxGlobal; gc%nyGlobal = nyGlobal; gc%nzGlobal = nzGlobal
    gc%wk0 = max(1, idz*nz + 1 - 12); gc%wk1 = min(nzGlobal, idz*nz + nz + 12)
    gc%idz = idz; gc%npz = npz
    gc%iper = iper; gc%jper = jper; gc%kper = kper
    gc%device = -1
    gc%theta = theta
    call mflbm_check(mflbm_geometry_preprocess(gc, c_loc(walls_global(1, 1, gc%wk0)), ps, ns, pf, nf, gs, gf), &
                     "mflbm_geometry_preprocess")
    num_solid_boundary = ns; num_fluid_boundary = nf
    num_solid_boundary_global = int(gs); num_fluid_boundary_global = int(gf)
    call c_f_pointer(ps, s, [max(ns, 1)]); call c_f_pointer(pf, f, [max(nf, 1)])
    if (allocated(solid_boundary_nodes)) deallocate(solid_boundary_nodes)
    if (allocated(fluid_boundary_nodes)) deallocate(fluid_boundary_nodes)
    allocate(solid_boundary_nodes(max(ns, 1)), fluid_boundary_nodes(max(nf, 1)))
    if (ns > 0) solid_boundary_nodes(1:ns) = s(1:ns)
    if (nf > 0) fluid_boundary_nodes(1:nf) = f(1:nf)
    call mflbm_geometry_free(ps); call mflbm_geometry_free(pf)
end subroutine geometry_preprocessing_new
