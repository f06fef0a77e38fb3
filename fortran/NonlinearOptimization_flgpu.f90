!Fortran host side of libflgpu.so: a module with the reference's own procedure names, argument order,
!optional arguments and defaults for the hot path (reference: source/NonlinearOptimization.f90
!LBFGS 398-400, ConjugateGradient 193-195, SteepestDescent 55-56), so that a program
!written against `use NonlinearOptimization` switches to the B200 path by compiling against this
!module and linking -lflgpu instead of -lFL.  Host control flow stays in the library's C++ driver;
!this file only maps Fortran OPTIONAL / LOGICAL / CHARACTER(*) onto the iso_c_binding C-ABI of
!include/flgpu.h (absent optional -> C_NULL_PTR, logical -> 4-byte integer, Method -> char* + length).
!
!NOT COMPILED IN THIS REPOSITORY'S IMAGE: no Fortran compiler exists there (SURVEY.md F1).  Build with
!    gfortran -c NonlinearOptimization_flgpu.f90 && gfortran prog.f90 NonlinearOptimization_flgpu.o -lflgpu
!The callbacks keep the reference's interface (f90:33-38).  By default they are HOST callbacks, as in the
!reference: the library stages x / fdx through pinned host buffers around every call, so unmodified code
!runs as it is (slow: every evaluation crosses PCIe).  A callback that launches CUDA kernels itself calls
!flgpu_set_callback_space(1) first and then receives DEVICE pointers on flgpu_current_stream().
!Written to the letter of Fortran 2008 where this image could not check it with a compiler: explicit
!interfaces for the callback dummies (c_funloc of a dummy procedure), character(*) Method copied into a
!character(kind=c_char) array (c_loc of a len /= 1 character is not interoperable in F2003).
module NonlinearOptimization_flgpu
    use iso_c_binding
    implicit none

    abstract interface
        !the reference's callback contract (f90:33-38)
        subroutine flgpu_f_iface(fx, x, dim)
            integer, intent(in) :: dim
            real*8, intent(out) :: fx
            real*8, dimension(dim), intent(in) :: x
        end subroutine flgpu_f_iface
        subroutine flgpu_fd_iface(fdx, x, dim)
            integer, intent(in) :: dim
            real*8, dimension(dim), intent(out) :: fdx
            real*8, dimension(dim), intent(in) :: x
        end subroutine flgpu_fd_iface
        integer function flgpu_f_fd_iface(fx, fdx, x, dim)
            integer, intent(in) :: dim
            real*8, intent(out) :: fx
            real*8, dimension(dim), intent(out) :: fdx
            real*8, dimension(dim), intent(in) :: x
        end function flgpu_f_fd_iface
    end interface

    interface
        !gfortran-mangled entry points exported by libflgpu.so; all arguments by reference
        subroutine flgpu_lbfgs_ref(f, fd, x, dim, Memory, f_fd, Strong, Warning, MaxIteration, Precision, &
                MinStepLength, WolfeConst1, WolfeConst2, Increment) bind(C, name='__nonlinearoptimization_MOD_lbfgs')
            import :: c_funptr, c_ptr, c_int, c_double
            type(c_funptr), value :: f, fd, f_fd
            real(c_double), dimension(*), intent(inout) :: x
            integer(c_int), intent(in) :: dim
            type(c_ptr), value :: Memory, Strong, Warning, MaxIteration, Precision, MinStepLength, &
                WolfeConst1, WolfeConst2, Increment
        end subroutine flgpu_lbfgs_ref
        subroutine flgpu_sd_ref(f, fd, x, dim, f_fd, Strong, Warning, MaxIteration, Precision, &
                MinStepLength, WolfeConst1, WolfeConst2, Increment) bind(C, name='__nonlinearoptimization_MOD_steepestdescent')
            import :: c_funptr, c_ptr, c_int, c_double
            type(c_funptr), value :: f, fd, f_fd
            real(c_double), dimension(*), intent(inout) :: x
            integer(c_int), intent(in) :: dim
            type(c_ptr), value :: Strong, Warning, MaxIteration, Precision, MinStepLength, &
                WolfeConst1, WolfeConst2, Increment
        end subroutine flgpu_sd_ref
        subroutine flgpu_cg_ref(f, fd, x, dim, Method, f_fd, Strong, Warning, MaxIteration, Precision, &
                MinStepLength, WolfeConst1, WolfeConst2, Increment, len_Method) &
                bind(C, name='__nonlinearoptimization_MOD_conjugategradient')
            import :: c_funptr, c_ptr, c_int, c_double
            type(c_funptr), value :: f, fd, f_fd
            real(c_double), dimension(*), intent(inout) :: x
            integer(c_int), intent(in) :: dim
            type(c_ptr), value :: Method, Strong, Warning, MaxIteration, Precision, MinStepLength, &
                WolfeConst1, WolfeConst2, Increment
            integer(c_int), value :: len_Method
        end subroutine flgpu_cg_ref
        subroutine flgpu_set_callback_space(space) bind(C, name='flgpu_set_callback_space')
            import :: c_int
            integer(c_int), value :: space
        end subroutine flgpu_set_callback_space
        subroutine flgpu_set_x_space(space) bind(C, name='flgpu_set_x_space')
            import :: c_int
            integer(c_int), value :: space
        end subroutine flgpu_set_x_space
        !0 = the reference's line searchers (default), 1 = FLGPU_LS_FAST (not a reference routine), -1 = unset
        subroutine flgpu_set_line_search(policy) bind(C, name='flgpu_set_line_search')
            import :: c_int
            integer(c_int), value :: policy
        end subroutine flgpu_set_line_search
        function flgpu_current_stream() bind(C, name='flgpu_current_stream') result(stream)
            import :: c_ptr
            type(c_ptr) :: stream
        end function flgpu_current_stream
    end interface

contains

    !Same dummy-argument list as the reference's LBFGS (f90:398-400)
    subroutine LBFGS(f, fd, x, dim, Memory, f_fd, Strong, Warning, MaxIteration, Precision, MinStepLength, &
            WolfeConst1, WolfeConst2, Increment)
        procedure(flgpu_f_iface) :: f
        procedure(flgpu_fd_iface) :: fd
        procedure(flgpu_f_fd_iface), optional :: f_fd
        integer, intent(in) :: dim
        real*8, dimension(dim), intent(inout), target :: x
        integer, intent(in), optional, target :: Memory, MaxIteration
        logical, intent(in), optional :: Strong, Warning
        real*8, intent(in), optional, target :: Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment
        integer(c_int), target :: istrong, iwarning
        type(c_funptr) :: cffd
        type(c_ptr) :: pstrong, pwarning
        cffd = c_null_funptr; if (present(f_fd)) cffd = c_funloc(f_fd)
        call logical_arg(Strong, istrong, pstrong); call logical_arg(Warning, iwarning, pwarning)
        call flgpu_lbfgs_ref(c_funloc(f), c_funloc(fd), x, dim, iptr(Memory), cffd, pstrong, pwarning, &
            iptr(MaxIteration), dptr(Precision), dptr(MinStepLength), dptr(WolfeConst1), dptr(WolfeConst2), &
            dptr(Increment))
    end subroutine LBFGS

    !Same dummy-argument list as the reference's SteepestDescent (f90:55-56)
    subroutine SteepestDescent(f, fd, x, dim, f_fd, Strong, Warning, MaxIteration, Precision, MinStepLength, &
            WolfeConst1, WolfeConst2, Increment)
        procedure(flgpu_f_iface) :: f
        procedure(flgpu_fd_iface) :: fd
        procedure(flgpu_f_fd_iface), optional :: f_fd
        integer, intent(in) :: dim
        real*8, dimension(dim), intent(inout), target :: x
        integer, intent(in), optional, target :: MaxIteration
        logical, intent(in), optional :: Strong, Warning
        real*8, intent(in), optional, target :: Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment
        integer(c_int), target :: istrong, iwarning
        type(c_funptr) :: cffd
        type(c_ptr) :: pstrong, pwarning
        cffd = c_null_funptr; if (present(f_fd)) cffd = c_funloc(f_fd)
        call logical_arg(Strong, istrong, pstrong); call logical_arg(Warning, iwarning, pwarning)
        call flgpu_sd_ref(c_funloc(f), c_funloc(fd), x, dim, cffd, pstrong, pwarning, iptr(MaxIteration), &
            dptr(Precision), dptr(MinStepLength), dptr(WolfeConst1), dptr(WolfeConst2), dptr(Increment))
    end subroutine SteepestDescent

    !Same dummy-argument list as the reference's ConjugateGradient (f90:193-195)
    subroutine ConjugateGradient(f, fd, x, dim, Method, f_fd, Strong, Warning, MaxIteration, Precision, &
            MinStepLength, WolfeConst1, WolfeConst2, Increment)
        procedure(flgpu_f_iface) :: f
        procedure(flgpu_fd_iface) :: fd
        procedure(flgpu_f_fd_iface), optional :: f_fd
        integer, intent(in) :: dim
        real*8, dimension(dim), intent(inout), target :: x
        character(*), intent(in), optional :: Method          !as the reference (f90:201); 'DY' or 'PR'
        character(kind=c_char), dimension(32), target :: mbuf
        integer, intent(in), optional, target :: MaxIteration
        logical, intent(in), optional :: Strong, Warning
        real*8, intent(in), optional, target :: Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment
        integer(c_int), target :: istrong, iwarning
        type(c_funptr) :: cffd
        type(c_ptr) :: pstrong, pwarning, pmethod
        integer(c_int) :: lmethod
        integer :: ich
        cffd = c_null_funptr; if (present(f_fd)) cffd = c_funloc(f_fd)
        pmethod = c_null_ptr; lmethod = 0
        if (present(Method)) then
            lmethod = min(len(Method), 32)
            do ich = 1, lmethod
                mbuf(ich) = Method(ich:ich)
            end do
            pmethod = c_loc(mbuf)
        end if
        call logical_arg(Strong, istrong, pstrong); call logical_arg(Warning, iwarning, pwarning)
        call flgpu_cg_ref(c_funloc(f), c_funloc(fd), x, dim, pmethod, cffd, pstrong, pwarning, &
            iptr(MaxIteration), dptr(Precision), dptr(MinStepLength), dptr(WolfeConst1), dptr(WolfeConst2), &
            dptr(Increment), lmethod)
    end subroutine ConjugateGradient

    !absent optional -> C_NULL_PTR (what gfortran itself passes for an absent OPTIONAL dummy)
    function iptr(v) result(p)
        integer, intent(in), optional, target :: v
        type(c_ptr) :: p
        p = c_null_ptr; if (present(v)) p = c_loc(v)
    end function iptr
    function dptr(v) result(p)
        real*8, intent(in), optional, target :: v
        type(c_ptr) :: p
        p = c_null_ptr; if (present(v)) p = c_loc(v)
    end function dptr
    subroutine logical_arg(v, store, p)
        logical, intent(in), optional :: v
        integer(c_int), intent(out), target :: store
        type(c_ptr), intent(out) :: p
        p = c_null_ptr; store = 0
        if (present(v)) then
            if (v) store = 1
            p = c_loc(store)
        end if
    end subroutine logical_arg

end module NonlinearOptimization_flgpu
