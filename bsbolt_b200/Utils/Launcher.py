"""`bsbolt Align` launcher (reference: bsbolt/Utils/Launcher.py:41-115): same argv construction, same
printed summary; the aligner behind it is the in-process GPU library instead of a `bwa` subprocess."""
import datetime
import os
import time

from bsbolt_b200.Align.AlignReads import BisulfiteAlignmentAndProcessing

bwa_path = 'bsbolt_b200'  # element 0 of the command list; kept for interface parity, not executed


def build_alignment_command(arguments):
    """argv exactly as the reference builds it for `bwa mem` (Launcher.py:75-115)"""
    bsb_command_dict = {arg[0]: str(arg[1]) for arg in arguments._get_kwargs()}
    bwa_cmd = [bwa_path, 'mem', '-Y']
    if bsb_command_dict['UN'] == 'True':
        bwa_cmd.extend(['-z'])
    for arg in ['M', 'S', 'j', 'p']:
        if bsb_command_dict[arg] == 'True':
            bwa_cmd.append(f'-{arg}')
    for arg in ['A', 'B', 'D', 'E', 'L', 'T', 'U', 'W', 'c', 'd', 'k', 'm', 'r', 't', 'w', 'y']:
        bwa_cmd.extend([f'-{arg}', bsb_command_dict[arg]])
    if bsb_command_dict['H'] != 'None':
        bwa_cmd.extend(['-H', bsb_command_dict['H']])
    if bsb_command_dict['I'] != 'None':
        bwa_cmd.extend(['-I', bsb_command_dict['I']])
    if bsb_command_dict['INDEL']:
        bwa_cmd.extend(['-O', bsb_command_dict['INDEL']])
    if bsb_command_dict['XA']:
        bwa_cmd.extend(['-h', bsb_command_dict['XA']])
    bwa_cmd.extend(['-e', bsb_command_dict['SP']])
    bwa_cmd.extend(['-l', bsb_command_dict['CP']])
    bwa_cmd.extend(['-n', bsb_command_dict['CT']])
    bwa_cmd.extend(['-Z', bsb_command_dict['DR']])
    if bsb_command_dict.get('K', 'None') != 'None':
        bwa_cmd.extend(['-K', bsb_command_dict['K']])
    database = bsb_command_dict['DB']
    if not database.endswith('.fa'):
        if not database.endswith('/'):
            database = f'{database}/BSB_ref.fa'
        else:
            database = f'{database}BSB_ref.fa'
        assert os.path.exists(database), f'-DB {arguments.DB} does not exist, please index genome'
        assert os.path.exists(f'{database}.opac'), f'-DB {arguments.DB} not complete, please re-index genome'
    bwa_cmd.append(database)
    bwa_cmd.append(bsb_command_dict['F1'])
    assert os.path.exists(arguments.F1), f'-F1 {arguments.F1} does not exist, please check path'
    if bsb_command_dict['F2'] != 'None':
        bwa_cmd.append(bsb_command_dict['F2'])
        assert os.path.exists(arguments.F2), f'-F2 {arguments.F2} does not exist, please check path'
    return bwa_cmd


def process_mapping_statistics(mapping_dict):
    processed_list = []
    try:
        mappability = (mapping_dict['TotalAlignments'] - mapping_dict['Unaligned']) / mapping_dict['TotalAlignments']
    except ZeroDivisionError:
        mappability = 0.000
    processed_list.append(f'Total Reads: {mapping_dict["TotalReads"]}')
    processed_list.append(f'Mappability: {mappability * 100:.3f} %')
    processed_list.append('------------------------------')
    processed_list.append(f'Reads Mapped to Watson_C2T: {mapping_dict["W_C2T"]}')
    processed_list.append(f'Reads Mapped to Crick_C2T: {mapping_dict["C_C2T"]}')
    processed_list.append(f'Reads Mapped to Watson_G2A: {mapping_dict["W_G2A"]}')
    processed_list.append(f'Reads Mapped to Crick_G2A: {mapping_dict["C_G2A"]}')
    processed_list.append('------------------------------')
    processed_list.append(f'Unmapped Reads (Single / Paired Ends): {mapping_dict["Unaligned"]}')
    processed_list.append(f'Bisulfite Ambiguous: {mapping_dict["BSAmbiguous"]}')
    return '\n'.join(processed_list)


def align_bisulfite(bwa_cmd, output_path, output_threads, output_to_stdout, device=0):
    import sys
    start = time.time()
    # the reference prints these lines to stdout, which corrupts an -OS stream (Launcher.py:43); here
    # they go to stderr when SAM is on stdout
    info = sys.stderr if output_to_stdout else sys.stdout
    print(' '.join(bwa_cmd), file=info)
    bs_alignment = BisulfiteAlignmentAndProcessing(bwa_cmd, output_path, output_threads, output_to_stdout, device=device)
    bs_alignment.align_reads()
    alignment_time = datetime.timedelta(seconds=round(time.time() - start))
    print(f'Alignment Complete: Time {alignment_time}', file=info)
    print('------------------------------', file=info)
    print(process_mapping_statistics(bs_alignment.mapping_statistics), file=info)
    return bs_alignment


def launch_alignment(arguments):
    bwa_cmd = build_alignment_command(arguments)
    if arguments.O is None and not arguments.OS:
        raise FileNotFoundError("-O and -OS arguments empty, please specify output path")
    return align_bisulfite(bwa_cmd, arguments.O, arguments.OT, arguments.OS, device=getattr(arguments, 'GPU', [0]))


def launch_index(arguments):
    """`bsbolt Index` (reference: Launcher.py:18-38): same messages, same database directory; the index files come from the GPU builder"""
    from bsbolt_b200 import index_db
    if arguments.rrbs:
        print(f'Generating RRBS Database at {arguments.DB}: '
              f'lower bound {arguments.rrbs_lower}, upper bound {arguments.rrbs_upper}: '
              f'Cut Format {arguments.rrbs_cut_format}')
        print(index_db.restriction_sites(arguments.rrbs_cut_format))
        return index_db.build_rrbs_database(arguments.G, arguments.DB, device=arguments.GPU, lower_bound=arguments.rrbs_lower,
                                            upper_bound=arguments.rrbs_upper, cut_format=arguments.rrbs_cut_format, ignore_alt=arguments.IA)
    print(f'Generating WGBS Database at {arguments.DB}')
    return index_db.build_database(arguments.G, arguments.DB, device=arguments.GPU, ignore_alt=arguments.IA, mappable_regions=arguments.MR)


bsb_launch = {'Align': launch_alignment, 'Index': launch_index}
